#!/usr/bin/env python
"""Golden ELBO terms, gradients and Adamax updates of the AR(1) model, produced by the reference's OWN classes.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_step.py

TensorFlow 1.8 cannot run here, so `tests/golden/tf_shim.py` stands in for the TF library (about 30 ops, torch
float64, eager).  Everything that is the REFERENCE - AR.py's init_dist / IAF._create_flow / Flow_Stack /
VI_SSM.__init__ / _ELBO / build_flow (AR.py:24-234) and optimisers/adamax.py's AdamaxOptimizer - is imported from
/root/reference unmodified and executed: building the model evaluates it on the injected inputs, and
`model.train_step()` is `sess.run(self.train_step)`.

Injected (the hot path's inputs): the feed (time_feats, mask, shift - built by the oracle's
gather, which ar_golden.npz pins bit-exactly to the reference's own feed code), the base noise eps
(init_dist.sample, AR.py:32), the theta sample (theta_dist.sample, AR.py:117; its log-density is injected as 0:
the theta posterior is built in main() from tf.contrib bijectors and is outside this fixture), and the initial
values of the variables, handed out in creation order from the product's flat parameter blob - so the fixture
also pins the blob layout to TF's variable creation order.

Output: ar_step_golden.npz, two cases
  small_*  p=4, kernel_len=10, batch_dims=7, 2 flows, feat_window=3 on dat/AR_*.txt: everything in full
           (terms, path, gradient of -ELBO, Adamax-updated variables; the pre-training gradient as per-variable norms + leading entries)
  full_*   hyperparameters.txt shapes (kernel_len=50, batch_dims=50, 3 flows, feat_window=10) at p=3: terms and path in
           full, gradients as per-variable norms + leading entries (the blob has 431 706 entries)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tf_shim  # noqa: E402

tf_shim.install()
sys.path.insert(0, REF)          # `from optimisers.adamax import AdamaxOptimizer` resolves to the reference's file
_spec = importlib.util.spec_from_file_location("reference_AR", os.path.join(REF, "AR.py"))
RAR = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(RAR)    # module level: seeds, a (mock) session; main() is not run

from oracle import nma_oracle as O  # noqa: E402
from viforssms_b200 import feed  # noqa: E402
from viforssms_b200.config import ar_config, param_layout  # noqa: E402

PRIORS = [(0.0, 10.0), (0.0, 10.0), (0.0, 10.0)]      # main.py / hyperparameters.txt
X0, OBS_STD, T = 10.0, 1.0, 5000


def load_dat():
    d = os.path.join(REF, "dat")
    return (np.loadtxt(os.path.join(d, "AR_obs_partial.txt"), np.float32),
            np.loadtxt(os.path.join(d, "AR_obs_binary.txt"), np.float32),
            np.loadtxt(os.path.join(d, "AR_time_till.txt"), np.float32))


def run_case(p, K, B, F, fw, seed, lr=1e-3, clip=2.5e8):
    obs, obs_bin, tt = load_dat()
    cfg = ar_config(p=p, K=K, B=B, F=F, H=1, feat_window=fw, T=T, obs_std=OBS_STD, x0=X0)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    params = O.glorot_init(layout, n, g, torch.float32)
    for name, (off, shape) in layout.items():           # non-zero biases, bounded activations (as the GPU tests do)
        k = int(np.prod(shape))
        if name.endswith(".b"):
            params[off:off + k] = 0.05 * torch.randn(k, generator=g)
    for i in range(F):
        off, shape = layout[f"f{i}.feat0.w"]
        params[off:off + int(np.prod(shape))].reshape(shape)[fw + 1, :] *= 10.0 / T
    eps = torch.randn(p, cfg.L0, generator=g)
    theta = torch.stack([torch.randn(p, generator=g) * 0.5 + 4.0, torch.randn(p, generator=g) * 0.1 + 0.5,
                         torch.randn(p, generator=g) * 0.2 + 1.0], dim=1).float()
    idx = feed.sample_indices(T, B, p, rs)
    pads = O.pad_series_ar(obs, obs_bin, tt, X0, T, F, K, fw)
    tf64, mask, shift = O.gather_feed_ar(pads, idx, cfg.L0, B)
    tf32 = tf64.astype(np.float32)                       # the feed_dict cast to DTYPE = float32 (AR.py:10,151-152)

    st = tf_shim.STATE
    st.__init__()
    st.placeholders = [tf32, mask, shift]                # AR.py:151-159, in creation order
    st.samples = [eps.numpy()]                           # init_dist.sample(p), AR.py:32 (one base sample per stack)
    st.blob = params.double()
    theta_dist = tf_shim.InjectedDistribution(theta.double().numpy(), np.zeros(p))

    model = RAR.VI_SSM(obs, OBS_STD, X0, theta_dist, PRIORS, T, p, K, B, [50, 50, 50], F, fw, obs_bin, tt,
                       pre_train=False, learn_rate=lr, grad_clip=clip)
    model.build_flow()
    assert st.cursor == n, "the reference created %d parameters, the product's layout has %d" % (st.cursor, n)
    # the product's layout lists variables in the same order with the same shapes
    for (name, (off, shape)), got in zip(sorted(layout.items(), key=lambda kv: kv[1][0]), st.var_shapes):
        got = got[1:] if (len(got) == 3 and got[0] == 1) else got      # kernel_size-1 conv1d kernels are [1, in, out]
        assert tuple(shape) == tuple(got), (name, shape, got)

    th = model.theta
    scale = float(T) / B
    # d/dtheta of the part of -ELBO the device path owns (the prior and log q(theta) are added on the host)
    dev_obj = -(scale * (model.sde_loss - model.lf_log_prob + model.obs_loss)).sum()
    g_theta = torch.autograd.grad(dev_obj, th, retain_graph=True)[0]
    elbo, _, _ = model._ELBO()
    grads = [gg for gg in model.gradients]               # clipped gradients of -ELBO (clip far above the norm: identity)
    flat_grad = torch.cat([gg.reshape(-1) for gg in grads])
    gnorm = tf_shim.global_norm(grads)
    # pre-training objective (AR.py:201-202): gradients of -obs_loss through the reference's optimizer API
    pre = RAR.AdamaxOptimizer(learning_rate=1e-3, beta1=0.9).compute_gradients(-model.obs_loss)
    flat_pre = torch.cat([(gg if gg is not None else torch.zeros_like(v.value)).reshape(-1) for gg, v in pre])

    out = {
        "hyper": np.array([p, K, B, F, fw, T, seed]), "idx": idx, "eps": eps.numpy(), "theta": theta.numpy(),
        "params": params.numpy(), "lr_clip": np.array([lr, clip]),
        "sde": model.sde_loss.detach().numpy(), "obs": model.obs_loss.detach().numpy(),
        "logq": model.lf_log_prob.detach().numpy(), "lf_sample": model.lf_sample.detach().numpy(),
        "elbo": elbo.detach().numpy(), "grad_theta": g_theta.numpy(), "global_norm": np.array(float(gnorm)),
    }
    # one Adamax step through the reference's optimizer (AR.py:226-234 -> optimisers/adamax.py:42-58)
    model.train_step()
    new_params = torch.cat([v.value.detach().reshape(-1) for v in st.variables])
    return out, flat_grad.detach(), flat_pre.detach(), new_params, layout


def main():
    golden = {}
    out, grad, pre, newp, layout = run_case(p=4, K=10, B=7, F=2, fw=3, seed=11)
    sha = lambda a: __import__("hashlib").sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for k, v in out.items():
        if k != "params":                                 # regenerated from the seed by the test, checked by hash
            golden["small_" + k] = v
    golden["small_params_sha_f32"] = np.array(sha(out["params"]))
    golden["small_grad"] = grad.numpy()                   # float64: the oracle is held to 1e-9 against it
    golden["small_params_after"] = newp.numpy().astype(np.float32)
    names = [kv[0] for kv in sorted(layout.items(), key=lambda kv: kv[1][0])]
    golden["small_var_names"] = np.array(names)
    golden["small_grad_pretrain_norms"] = np.array([float(pre[layout[nm][0]:layout[nm][0] + int(np.prod(layout[nm][1]))].norm())
                                                    for nm in names])
    golden["small_grad_pretrain_heads"] = np.concatenate([pre[layout[nm][0]:layout[nm][0] + min(int(np.prod(layout[nm][1])), 16)].numpy()
                                                          for nm in names])

    out, grad, pre, newp, layout = run_case(p=3, K=50, B=50, F=3, fw=10, seed=12)
    for k, v in out.items():
        if k != "params":                                 # 431 706 values: regenerated from the seed by the test
            golden["full_" + k] = v
    names = [kv[0] for kv in sorted(layout.items(), key=lambda kv: kv[1][0])]
    norms, heads = [], []
    for nm in names:
        off, shape = layout[nm]
        k = int(np.prod(shape))
        norms.append(float(grad[off:off + k].norm()))
        heads.append(grad[off:off + min(k, 16)].numpy())
    golden["full_var_names"] = np.array(names)
    golden["full_grad_norms"] = np.array(norms)
    golden["full_grad_heads"] = np.concatenate(heads)
    golden["full_params_sha_f32"] = np.array(sha(out["params"]))
    golden["full_params_after_head"] = newp[:4096].numpy()
    golden["full_pretrain_grad_norm"] = np.array(float(pre.norm()))
    path = os.path.join(HERE, "ar_step_golden.npz")
    np.savez_compressed(path, **golden)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("small_sde", "small_obs", "small_logq", "small_elbo", "small_global_norm", "full_elbo", "full_global_norm"):
        print(k, golden[k])


if __name__ == "__main__":
    main()
