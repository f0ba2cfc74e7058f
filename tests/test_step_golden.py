"""The oracle's flow / ELBO / gradient / Adamax restatement against values produced by the reference's OWN classes
(tests/golden/ar_step_golden.npz, made by tests/golden/make_golden_step.py: AR.py's init_dist, IAF._create_flow,
Flow_Stack, VI_SSM._ELBO / build_flow and optimisers/adamax.py's AdamaxOptimizer, imported unmodified from the
reference and executed over tests/golden/tf_shim.py, a torch float64 stand-in for the ~30 TensorFlow-1.8 library
ops they call).  What is pinned: the composition - slices, terms, signs, the T / batch_dims scale, variable
creation order, the optimiser's slot arithmetic.  What is not: TensorFlow's own op kernels (restated in the shim
from their documented behaviour) and the theta posterior (A11), whose sample and log-density are injected.

CPU only; the GPU leg is tests/test_gpu_step_golden.py."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import nma_oracle as O
from viforssms_b200 import feed
from viforssms_b200.config import ar_config, param_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRIORS = [(0.0, 10.0), (0.0, 10.0), (0.0, 10.0)]


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "ar_step_golden.npz"))


def case_inputs(G, case):
    """Rebuilds the inputs of one golden case exactly as make_golden_step.run_case did (same seeds, same code path),
    and checks them against the stored eps / theta / idx and the hash of the parameter blob."""
    p, K, B, F, fw, T, seed = (int(v) for v in G[case + "_hyper"])
    d = os.path.join(ROOT, "dat")
    obs = np.loadtxt(os.path.join(d, "AR_obs_partial.txt"), np.float32)
    obs_bin = np.loadtxt(os.path.join(d, "AR_obs_binary.txt"), np.float32)
    tt = np.loadtxt(os.path.join(d, "AR_time_till.txt"), np.float32)
    cfg = ar_config(p=p, K=K, B=B, F=F, H=1, feat_window=fw, T=T, obs_std=1.0, x0=10.0)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    params = O.glorot_init(layout, n, g, torch.float32)
    for name, (off, shape) in layout.items():
        k = int(np.prod(shape))
        if name.endswith(".b"):
            params[off:off + k] = 0.05 * torch.randn(k, generator=g)
    for i in range(F):
        off, shape = layout[f"f{i}.feat0.w"]
        params[off:off + int(np.prod(shape))].reshape(shape)[fw + 1, :] *= 10.0 / T
    sha = hashlib.sha256(np.ascontiguousarray(params.numpy()).tobytes()).hexdigest()
    assert sha == str(G[case + "_params_sha_f32"]), "torch's CPU generator no longer reproduces the fixture's weights"
    eps = torch.from_numpy(G[case + "_eps"])
    theta = torch.from_numpy(G[case + "_theta"])
    idx = G[case + "_idx"]
    pads = O.pad_series_ar(obs, obs_bin, tt, 10.0, T, F, K, fw)
    tf64, _, _ = O.gather_feed_ar(pads, idx, cfg.L0, B)
    tf32 = torch.from_numpy(tf64.astype(np.float32))
    arrays = feed.ar_base_arrays(obs, obs_bin, tt, T, F, K, fw)
    return cfg, layout, n, params, eps, theta, idx, tf32, arrays


def _close(got, want, rtol):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.abs(got - want).max() <= rtol * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("case", ["small", "full"])
def test_oracle_terms_and_path_match_the_reference_classes(G, case):
    cfg, layout, n, params, eps, theta, idx, tf32, _ = case_inputs(G, case)
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf32.double())
    t = ref["terms"].numpy()
    assert _close(t[:, 0], G[case + "_sde"], 1e-10)           # VI_SSM.sde_loss   (AR.py:174-176)
    assert _close(t[:, 1], G[case + "_obs"], 1e-10)           # VI_SSM.obs_loss   (AR.py:169-170)
    assert _close(t[:, 2], G[case + "_logq"], 1e-10)          # lf_log_prob       (AR.py:33-34,84-88)
    assert _close(ref["lf"].numpy(), G[case + "_lf_sample"], 1e-10)
    # ELBO per row (AR.py:184-185) with log q(theta) injected as 0: scale (sde - logq + obs) + log prior(theta)
    th = theta.double().numpy()
    prior = sum(-0.5 * ((th[:, k] - m) / s) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(s) for k, (m, s) in enumerate(PRIORS))
    elbo = cfg.scale * (t[:, 0] - t[:, 2] + t[:, 1]) + prior
    assert _close(elbo, G[case + "_elbo"], 1e-10)


def test_oracle_gradients_match_the_reference_optimizer(G):
    """opt.compute_gradients(-loss) of AR.py:228-229 (sum over rows), every entry of the 83 k-parameter blob."""
    cfg, layout, n, params, eps, theta, idx, tf32, _ = case_inputs(G, "small")
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf32.double())
    want = G["small_grad"]
    got = ref["grad_params"].numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    for name, (off, shape) in layout.items():                 # and per variable, so that small tensors count too
        k = int(np.prod(shape))
        assert np.linalg.norm(got[off:off + k] - want[off:off + k]) <= 1e-9 * max(np.linalg.norm(want[off:off + k]), 1e-12), name
    assert _close(ref["grad_theta"].numpy(), G["small_grad_theta"], 1e-10)
    assert abs(float(ref["grad_params"].norm()) - float(G["small_global_norm"])) <= 1e-10 * float(G["small_global_norm"])


def test_oracle_gradients_full_shape(G):
    """hyperparameters.txt shapes (kernel_len 50, batch_dims 50, 3 flows): per-variable norms and leading entries."""
    cfg, layout, n, params, eps, theta, idx, tf32, _ = case_inputs(G, "full")
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf32.double())
    got = ref["grad_params"].numpy()
    names = [str(s) for s in G["full_var_names"]]
    assert names == [kv[0] for kv in sorted(layout.items(), key=lambda kv: kv[1][0])]
    heads = G["full_grad_heads"]
    pos = 0
    for nm, want_norm in zip(names, G["full_grad_norms"]):
        off, shape = layout[nm]
        k = int(np.prod(shape))
        assert abs(np.linalg.norm(got[off:off + k]) - want_norm) <= 1e-9 * max(want_norm, 1e-12), nm
        h = min(k, 16)
        assert np.abs(got[off:off + h] - heads[pos:pos + h]).max() <= 1e-9 * max(np.abs(heads[pos:pos + h]).max(), 1e-12), nm
        pos += h
    assert _close(ref["grad_theta"].numpy(), G["full_grad_theta"], 1e-10)


def test_pretrain_objective_gradient(G):
    """minimize(-obs_loss) of AR.py:201-202 = objective 1."""
    cfg, layout, n, params, eps, theta, idx, tf32, _ = case_inputs(G, "small")
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf32.double(), obj=1)
    got = ref["grad_params"].numpy()
    heads = G["small_grad_pretrain_heads"]
    pos = 0
    for nm, want_norm in zip([str(s) for s in G["small_var_names"]], G["small_grad_pretrain_norms"]):
        off, shape = layout[nm]
        k = int(np.prod(shape))
        assert abs(np.linalg.norm(got[off:off + k]) - want_norm) <= 1e-9 * max(want_norm, 1e-9), nm
        h = min(k, 16)
        assert np.abs(got[off:off + h] - heads[pos:pos + h]).max() <= 1e-9 * max(np.abs(heads[pos:pos + h]).max(), 1e-9), nm
        pos += h


def test_adamax_restatement_matches_the_reference_optimizer(G):
    """One apply_gradients of the reference's AdamaxOptimizer (optimisers/adamax.py:42-58) after
    clip_by_global_norm (AR.py:230-234), from zero slots: the oracle's adamax_step on the golden gradient."""
    cfg, layout, n, params, eps, theta, idx, tf32, _ = case_inputs(G, "small")
    lr, clip = (float(v) for v in G["small_lr_clip"])
    g = torch.from_numpy(G["small_grad"])
    gn = float(G["small_global_norm"])
    w, m, v = O.adamax_step(params.double(), g, torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64),
                            lr, 0.95, clip=(clip, gn))
    want = G["small_params_after"].astype(np.float64)           # stored as float32
    assert np.abs(w.numpy() - want).max() <= 2e-7 * max(1.0, np.abs(want).max())
    # the update itself (one part in 1e3 of the weights): lr * v / m with v = (1 - b1) g, m = max(1e-8, |g|)
    upd_got = (w - params.double()).numpy()
    upd_want = want - params.double().numpy()
    big = np.abs(G["small_grad"]) > 1e-3
    assert np.abs(upd_got[big] - upd_want[big]).max() <= 1e-4 * lr
