#!/usr/bin/env python
"""Golden ELBO terms, paths and gradients of the FitzHugh-Nagumo and stochastic-volatility models, produced by the
reference's OWN classes (fitz_nag_NVP.py:25-448, SV_dense.py:25-402) executed over tests/golden/tf_shim.py.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_step_models.py

Both scripts are module-level programs: everything above their line "########### setting up the model ###########"
defines the classes (init_dist, IAF, Permute, Flow_Stack, VI_SSM), everything below loads data files the reference
does not ship and runs.  We exec the class section verbatim - the source text is read from the reference file and
cut at that marker, nothing is edited - in a namespace that then receives the few module-level names the classes
read as globals (p, no_flows, network_dims), build VI_SSM with small shapes on a synthetic series, and store what
the model's tensors evaluate to.  Feed, base noise, the theta sample and the initial variable values are injected
exactly as in make_golden_step.py; the feed comes from the oracle's gather, pinned bit-exactly to the scripts' own
feed code by fhn_golden.npz / sv_golden.npz.

Output: models_step_golden.npz (fhn_*, sv_*, lvr_* = lotka_volterra_partial.py): inputs, per-row terms, path, gradient of -ELBO w.r.t. theta, and the
gradient w.r.t. every variable as per-variable norms + leading entries.
"""
import hashlib
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
sys.path.insert(0, REF)

from oracle import nma_oracle as O  # noqa: E402
from viforssms_b200.config import fhn_config, param_layout, sv_config  # noqa: E402

MARK = "########### setting up the model ###########"
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def class_section(script, **module_globals):
    src = open(os.path.join(REF, script)).read()
    ns = {"__name__": "reference_" + script[:-3]}
    exec(compile(src[:src.index(MARK)], os.path.join(REF, script), "exec"), ns)
    ns.update(module_globals)
    return ns


def init_params(cfg, layout, n, g, time_row, T):
    params = O.glorot_init(layout, n, g, torch.float32)
    for name, (off, shape) in layout.items():
        k = int(np.prod(shape))
        if name.endswith(".b") or name.endswith(".beta"):
            params[off:off + k] = 0.05 * torch.randn(k, generator=g)
        if name.endswith(".gamma"):
            params[off:off + k] = 1.0 + 0.1 * torch.randn(k, generator=g)
    if time_row is not None:
        for i in range(cfg.F):
            off, shape = layout[f"f{i}.feat0.w"]
            params[off:off + shape[0] * shape[1]].reshape(shape)[time_row, :] *= 10.0 / T
    return params


def grads_summary(flat, layout, prefix, golden):
    names = [kv[0] for kv in sorted(layout.items(), key=lambda kv: kv[1][0])]
    golden[prefix + "var_names"] = np.array(names)
    golden[prefix + "grad_norms"] = np.array([float(flat[layout[nm][0]:layout[nm][0] + int(np.prod(layout[nm][1]))].norm())
                                              for nm in names])
    golden[prefix + "grad_heads"] = np.concatenate(
        [flat[layout[nm][0]:layout[nm][0] + min(int(np.prod(layout[nm][1])), 16)].numpy() for nm in names])
    golden[prefix + "global_norm"] = np.array(float(flat.norm()))


def check_layout(st, layout, n):
    assert st.cursor == n, "the reference created %d parameters, the product's layout has %d" % (st.cursor, n)
    for (name, (off, shape)), got in zip(sorted(layout.items(), key=lambda kv: kv[1][0]), st.var_shapes):
        got = got[1:] if (len(got) == 3 and got[0] == 1) else got
        assert tuple(shape) == tuple(got), (name, shape, got)


def fhn(golden):
    p, K, B, F, fw, target_dims, dt, seed = 5, 6, 5, 3, 3, 240, 0.1, 21
    T = target_dims * dt
    cfg = fhn_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=target_dims, dt=dt)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    obs = rs.normal(0.5, 1.0, size=(2, target_dims))
    obs_bin = (rs.uniform(size=(2, target_dims)) < 0.3).astype(np.float64)
    obs = obs * obs_bin
    tt = rs.uniform(0.0, 1.0, size=(2, target_dims)).round(1)
    x0 = np.array([2.0, 3.0])
    pads = O.pad_series_fhn(obs, tt, x0, dt, T, target_dims, F, K, fw)
    idx = rs.choice(np.arange(0, target_dims, B), size=p, replace=False)
    idx[0] = 0
    tf64, mask, shift, bin_feed = O.gather_feed_fhn(pads, obs_bin, idx, cfg.L0, B)
    params = init_params(cfg, layout, n, g, None, T)
    eps = torch.randn(p, cfg.L0, generator=g)
    theta = torch.stack([torch.randn(p, generator=g) * 0.2 + 0.7, torch.randn(p, generator=g) * 0.2 + 1.0,
                         torch.randn(p, generator=g) * 0.2 + 1.5, torch.randn(p, generator=g) * 0.2 - 0.7,
                         torch.randn(p, generator=g) * 0.2 - 1.2], dim=1).float()
    priors = [(0.0, 10.0)] * 5
    network_dims = [50] * 5

    ns = class_section("fitz_nag_NVP.py", p=p, no_flows=F, network_dims=network_dims)
    st = tf_shim.STATE
    st.__init__()
    f32 = lambda a: np.asarray(a).astype(np.float32)
    st.placeholders = [np.ones(1), f32(tf64), f32(mask), f32(shift), f32(bin_feed)]   # fitz_nag_NVP.py:178,212-222
    st.samples = [eps.numpy()]
    st.blob = params.double()
    theta_dist = tf_shim.InjectedDistribution(theta.double().numpy(), np.zeros(p))
    model = ns["VI_SSM"](obs.astype(np.float32), obs_bin.astype(np.float32), tt.astype(np.float32), x0, theta_dist,
                         priors, dt, T, p, K, B, network_dims, target_dims, F, fw, learn_rate=1e-4, pre_train=False)
    model.build_flow()
    check_layout(st, layout, n)
    scale = float(target_dims) / B
    dev_obj = -(scale * (model.sde_loss - model.lf_log_prob + model.obs_loss)).sum()
    g_theta = torch.autograd.grad(dev_obj, model.theta, retain_graph=True)[0]
    gv = ns["AdamaxOptimizer"](learning_rate=1e-4, beta1=0.95).compute_gradients(-model.loss)
    flat = torch.cat([gg.reshape(-1) for gg, _ in gv]).detach()
    # the pre-training objective (lf_sample - 0)^2 of fitz_nag_NVP.py:288-289
    gv2 = ns["AdamaxOptimizer"](learning_rate=1e-3, beta1=0.9).compute_gradients((model.lf_sample - 0.0) ** 2)
    flat2 = torch.cat([(gg if gg is not None else torch.zeros_like(v.value)).reshape(-1) for gg, v in gv2]).detach()
    golden.update({
        "fhn_hyper": np.array([p, K, B, F, fw, target_dims, seed]), "fhn_dt": np.array(dt),
        "fhn_obs": obs, "fhn_obs_bin": obs_bin, "fhn_time_till": tt, "fhn_idx": idx.astype(np.int64),
        "fhn_eps": eps.numpy(), "fhn_theta": theta.numpy(), "fhn_params_sha_f32": np.array(sha(params.numpy())),
        "fhn_sde": model.sde_loss.detach().numpy(), "fhn_obs_lp": model.obs_loss.detach().numpy(),
        "fhn_logq": model.lf_log_prob.detach().numpy(), "fhn_lf_sample": model.lf_sample.detach().numpy(),
        "fhn_elbo": model.loss.detach().numpy(), "fhn_grad_theta": g_theta.numpy(),
    })
    grads_summary(flat, layout, "fhn_", golden)
    grads_summary(flat2, layout, "fhn_pre_", golden)
    print("fhn: sde", golden["fhn_sde"][:3], "global norm", float(flat.norm()))


def sv(golden):
    p, K, B, F, fw, N, dt, seed = 6, 10, 7, 3, 2, 330, 1.0, 22
    T = float(N)
    x0 = -8.5
    cfg = sv_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=x0)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    obs = (20.0 + np.cumsum(rs.standard_normal(N + 1) * 0.4)).astype(np.float32)      # a price series (SV.dat is float32)
    pads = O.pad_series_sv(obs, x0, dt, T, N, F, K, fw)
    idx = rs.choice(np.arange(0, N, B), size=p, replace=False).astype(np.int64)
    idx[0] = 0
    idx[-1] = ((N - B - 1) // B) * B
    tf64, mask, shift, dim_one = O.gather_feed_sv(pads, idx, cfg.L0, B)
    params = init_params(cfg, layout, n, g, fw, N)
    eps = torch.randn(p, cfg.L0, generator=g)
    theta = torch.stack([torch.randn(p, generator=g) * 0.001 + 0.001, torch.randn(p, generator=g) * 0.1 - 0.6,
                         torch.randn(p, generator=g) * 0.1 - 2.5, torch.randn(p, generator=g) * 0.1 - 0.7], dim=1).float()
    priors = [(0.0, 10.0)] * 4
    network_dims = [50] * 5

    ns = class_section("SV_dense.py", p=p, no_flows=F, network_dims=network_dims)
    st = tf_shim.STATE
    st.__init__()
    f32 = lambda a: np.asarray(a).astype(np.float32)
    st.placeholders = [f32(tf64), f32(mask), f32(shift), f32(dim_one)]                 # SV_dense.py:186-193
    st.samples = [eps.numpy()]
    st.blob = params.double()
    theta_dist = tf_shim.InjectedDistribution(theta.double().numpy(), np.zeros(p))
    model = ns["VI_SSM"](obs, x0, theta_dist, priors, dt, T, p, K, B, network_dims, N, F, fw, learn_rate=1e-3,
                         pre_train=False)
    model.build_flow()
    check_layout(st, layout, n)
    scale = float(N) / B
    dev_obj = -(scale * (model.sde_loss - model.lf_log_prob)).sum()
    g_theta = torch.autograd.grad(dev_obj, model.theta, retain_graph=True)[0]
    gv = ns["AdamaxOptimizer"](learning_rate=1e-3, beta1=0.95).compute_gradients(-model.loss)
    flat = torch.cat([gg.reshape(-1) for gg, _ in gv]).detach()
    gv2 = ns["AdamaxOptimizer"](learning_rate=1e-3, beta1=0.9).compute_gradients((model.lf_sample + 7.0) ** 2)
    flat2 = torch.cat([(gg if gg is not None else torch.zeros_like(v.value)).reshape(-1) for gg, v in gv2]).detach()
    golden.update({
        "sv_hyper": np.array([p, K, B, F, fw, N, seed]), "sv_dt": np.array(dt), "sv_x0": np.array(x0),
        "sv_obs": obs, "sv_idx": idx, "sv_eps": eps.numpy(), "sv_theta": theta.numpy(),
        "sv_params_sha_f32": np.array(sha(params.numpy())),
        "sv_sde": model.sde_loss.detach().numpy(), "sv_logq": model.lf_log_prob.detach().numpy(),
        "sv_lf_sample": model.lf_sample.detach().numpy(), "sv_elbo": model.loss.detach().numpy(),
        "sv_grad_theta": g_theta.numpy(),
    })
    grads_summary(flat, layout, "sv_", golden)
    grads_summary(flat2, layout, "sv_pre_", golden)
    print("sv: sde", golden["sv_sde"][:3], "global norm", float(flat.norm()))


def lvr(golden):
    """lotka_volterra_partial.py (learned theta) on the dat/LV_*.txt files the reference ships, at the script's own
    kernel_len / batch_dims / depth / look-ahead (:466-476), 6 rows."""
    p, K, B, F, fw, target_dims, dt, seed = 6, 20, 50, 3, 10, 500, 0.1, 23
    T = 50.0
    from viforssms_b200.config import lvr_config
    d = os.path.join(REF, "dat")
    obs = np.loadtxt(os.path.join(d, "LV_obs_partial.txt"), np.float32)
    obs_bin = np.loadtxt(os.path.join(d, "LV_obs_binary.txt"), np.float32)
    tt = np.loadtxt(os.path.join(d, "LV_time_till.txt"), np.float32)
    x0 = np.array([100.0, 100.0])
    cfg = lvr_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=target_dims, dt=dt, x0=x0)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    pads = O.pad_series_lvr(obs, tt, x0, dt, T, target_dims, F, K, fw)
    idx = rs.choice(np.arange(0, target_dims, B), size=p, replace=False)
    idx[0] = 0                                        # the row whose first state is pinned to x0
    tf64, mask, shift, bin_feed = O.gather_feed_fhn(pads, obs_bin, idx, cfg.L0, B)     # same gather code (:354-386)
    params = init_params(cfg, layout, n, g, None, T)
    for i in range(F):          # observations reach 325 and time_till 10: keep the first feature layer's output O(1)
        off, shape = layout[f"f{i}.feat0.w"]
        params[off:off + shape[0] * shape[1]] *= 0.02
    eps = torch.randn(p, cfg.L0, generator=g)
    # theta sample around the script's prior means log(rate / 10) (:476)
    theta = (torch.tensor([np.log(0.4428), np.log(0.0029), np.log(0.2957)]).float()[None, :]
             + 0.05 * torch.randn(p, 3, generator=g)).float()
    priors = [(np.log(4.428 / 10), 1e-4), (np.log(0.029 / 10), 1e-4), (np.log(2.957 / 10), 1e-4)]
    network_dims = [50] * 5

    ns = class_section("lotka_volterra_partial.py", p=p, no_flows=F, network_dims=network_dims, kernel_len=K)
    st = tf_shim.STATE
    st.__init__()
    f32 = lambda a: np.asarray(a).astype(np.float32)
    st.placeholders = [np.ones(1), f32(tf64), f32(mask), f32(shift), f32(bin_feed)]   # lotka_volterra_partial.py:180,207-217
    st.samples = [eps.numpy()]
    st.blob = params.double()
    theta_dist = tf_shim.InjectedDistribution(theta.double().numpy(), np.zeros(p))
    model = ns["VI_SSM"](obs, obs_bin, tt, x0, theta_dist, priors, dt, T, p, K, B, network_dims, target_dims, F, fw,
                         learn_rate=1e-3, pre_train=False)
    model.build_flow()
    check_layout(st, layout, n)
    scale = float(target_dims) / B
    dev_obj = -(scale * (model.sde_loss - model.lf_log_prob + model.obs_loss)).sum()
    g_theta = torch.autograd.grad(dev_obj, model.theta, retain_graph=True)[0]
    gv = ns["AdamaxOptimizer"](learning_rate=1e-3, beta1=0.95).compute_gradients(-model.loss)
    flat = torch.cat([gg.reshape(-1) for gg, _ in gv]).detach()
    gv2 = ns["AdamaxOptimizer"](learning_rate=1e-3, beta1=0.9).compute_gradients((model.lf_sample - 75) ** 2)   # :299-300
    flat2 = torch.cat([(gg if gg is not None else torch.zeros_like(v.value)).reshape(-1) for gg, v in gv2]).detach()
    golden.update({
        "lvr_hyper": np.array([p, K, B, F, fw, target_dims, seed]), "lvr_dt": np.array(dt), "lvr_idx": idx.astype(np.int64),
        "lvr_obs": obs, "lvr_obs_bin": obs_bin, "lvr_time_till": tt,     # the three 2 x 500 input series (dat/LV_*.txt)
        "lvr_eps": eps.numpy(), "lvr_theta": theta.numpy(), "lvr_params_sha_f32": np.array(sha(params.numpy())),
        "lvr_sde": model.sde_loss.detach().numpy(), "lvr_obs_lp": model.obs_loss.detach().numpy(),
        "lvr_logq": model.lf_log_prob.detach().numpy(), "lvr_lf_sample": model.lf_sample.detach().numpy(),
        "lvr_elbo": model.loss.detach().numpy(), "lvr_grad_theta": g_theta.numpy(),
    })
    grads_summary(flat, layout, "lvr_", golden)
    grads_summary(flat2, layout, "lvr_pre_", golden)
    print("lvr: sde", golden["lvr_sde"][:3], "obs", golden["lvr_obs_lp"][:3], "global norm", float(flat.norm()))


def lv_fixed(golden):
    """lotka_volterra_partial_batch_fix_theta.py (fixed theta) at p_val = 3 short series, under BOTH readings of how
    Softplus(event_ndims=2) reduces the log-determinant of the flattened [states, 2] matrix (tf_shim.EVENT_REDUCTION)."""
    from viforssms_b200.config import lv_config
    p, K, B, F, fw, dt, seed = 3, 4, 6, 2, 2, 0.2, 24
    N = p * B
    T = (N - 1) * dt
    cfg = lv_config(p=p, K=K, B=B, F=F, H=2, feat_window=fw, target_dims=B, dt=dt, x0=(11.0, 9.5))
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    t = np.arange(N, dtype=np.float64)
    obs = np.stack([12.0 + 4.0 * np.sin(0.3 * t) + 0.3 * rs.standard_normal(N),
                    9.0 + 3.0 * np.cos(0.3 * t) + 0.3 * rs.standard_normal(N)])
    obs_bin = (rs.uniform(size=(2, N)) < 0.7).astype(np.float64)
    obs = np.where(obs_bin > 0, obs, np.log1p(np.exp(-2.0)) + 1.0)
    tt = rs.uniform(0.0, 1.0, size=(2, N)).round(1)
    x0 = np.array(cfg.x0)
    x0_std = np.array([1.0, 1.0])
    pads = O.pad_series_lv(obs, tt, x0, dt, T, N, 1, F, K, fw)
    idx = np.arange(p, dtype=np.int64) * B
    tf64, mask, shift, bin_feed = O.gather_feed_lv(pads, obs_bin, idx, cfg.L0, B)
    params = init_params(cfg, layout, n, g, None, T)
    for i in range(F):
        off, _ = layout[f"f{i}.head.b"]
        params[off] = 3.0
    eps = torch.randn(p, cfg.L0, generator=g)
    theta_vals = np.log1p(np.exp([-1.0, -6.0, -1.0, -2.0]))          # the script's constants (:688-691), softplus'd
    network_dims = [50] * 4
    f32 = lambda a: np.asarray(a).astype(np.float32)
    golden.update({"lvf_hyper": np.array([p, K, B, F, fw, seed]), "lvf_dt": np.array(dt), "lvf_obs": obs,
                   "lvf_obs_bin": obs_bin, "lvf_time_till": tt, "lvf_eps": eps.numpy(), "lvf_theta": theta_vals,
                   "lvf_params_sha_f32": np.array(sha(params.numpy()))})
    for reading in ("per_state", "literal"):
        tf_shim.EVENT_REDUCTION = reading
        ns = class_section("lotka_volterra_partial_batch_fix_theta.py", p_val=p, no_flows=F, network_dims=network_dims,
                           kernel_len=K)
        st = tf_shim.STATE
        st.__init__()
        st.placeholders = [np.ones(1), f32(tf64), f32(mask), f32(shift), f32(bin_feed)]      # :197,234-247
        st.samples = [eps.numpy()]
        st.blob = params.double()
        # the script's `priors` ARE its theta values (:190); p_val series of target_dims = B steps each
        model = ns["VI_SSM"](obs, obs_bin, tt, x0, x0_std, list(theta_vals), dt, T, p, K, B, network_dims, B, F, fw,
                             learn_rate=1e-3, pre_train=False)
        model.build_flow()
        check_layout(st, layout, n)
        gv = ns["AdamaxOptimizer"](learning_rate=1e-3, beta1=0.95).compute_gradients(-model.loss)
        flat = torch.cat([gg.reshape(-1) for gg, _ in gv]).detach()
        pre = "lvf_%s_" % reading
        golden.update({pre + "sde": model.sde_loss.detach().numpy(), pre + "obs_lp": model.obs_loss.detach().numpy(),
                       pre + "logq": model.lf_log_prob.detach().numpy(),
                       pre + "lf_sample": model.lf_sample.detach().numpy(), pre + "elbo": model.loss.detach().numpy()})
        grads_summary(flat, layout, pre, golden)
        print("lv fixed theta (%s): sde" % reading, golden[pre + "sde"], "global norm", float(flat.norm()))
    tf_shim.EVENT_REDUCTION = "literal"


def lv_batch(golden):
    """lotka_volterra_partial_batch.py - the fixed-theta script's flow with a LEARNED theta (softplus-transformed posterior
    and prior, :198-200,358-365), the plain bivariate transition density (no bijector chain, :339-343) and p_val = 3 windows
    per iteration whose first p_val states are pinned (mask_vals, :237-240) - at short series."""
    from viforssms_b200.config import lv_config
    p, K, B, F, fw, dt, seed = 3, 4, 6, 2, 2, 0.2, 31
    N = p * B
    T = (B - 1) * dt                     # the script's T is the length of ONE series: target_dims = T / dt + 1 = B (:681-683)
    cfg = lv_config(p=p, K=K, B=B, F=F, H=2, feat_window=fw, target_dims=B, dt=dt, x0=(11.0, 9.5))
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    t = np.arange(N, dtype=np.float64)
    obs = np.stack([12.0 + 4.0 * np.sin(0.3 * t) + 0.3 * rs.standard_normal(N),
                    9.0 + 3.0 * np.cos(0.3 * t) + 0.3 * rs.standard_normal(N)])
    obs_bin = (rs.uniform(size=(2, N)) < 0.7).astype(np.float64)
    obs = np.where(obs_bin > 0, obs, np.log1p(np.exp(-2.0)) + 1.0)
    tt = rs.uniform(0.0, 1.0, size=(2, N)).round(1)
    x0 = np.array(cfg.x0)
    x0_std = np.array([1.0, 1.0])
    pads = O.pad_series_lv(obs, tt, x0, dt, T, B, p, F, K, fw)
    idx = np.arange(p, dtype=np.int64) * B
    tf64, mask, shift, bin_feed = O.gather_feed_lv(pads, obs_bin, idx, cfg.L0, B)
    params = init_params(cfg, layout, n, g, None, T)
    for i in range(F):
        off, _ = layout[f"f{i}.head.b"]
        params[off] = 3.0
    eps = torch.randn(p, cfg.L0, generator=g)
    theta = torch.tensor(np.log1p(np.exp([-1.0, -6.0, -1.0, -2.0]))).repeat(p, 1) * (1.0 + 0.1 * torch.randn(p, 4, generator=g).double())
    theta_lp = torch.randn(p, generator=g).double()
    priors = [(-1.0, np.sqrt(0.1)), (-6.0, np.sqrt(0.1)), (-1.0, np.sqrt(0.1)), (-2.0, np.sqrt(0.1))]     # :689-690
    network_dims = [50] * 4
    f32 = lambda a: np.asarray(a).astype(np.float32)
    tf_shim.EVENT_REDUCTION = "per_state"
    ns = class_section("lotka_volterra_partial_batch.py", p_val=p, no_flows=F, network_dims=network_dims, kernel_len=K)
    st = tf_shim.STATE
    st.__init__()
    st.placeholders = [np.ones(1), f32(tf64), f32(mask), f32(shift), f32(bin_feed)]
    st.samples = [eps.numpy()]
    st.blob = params.double()
    theta_dist = tf_shim.InjectedDistribution(theta.numpy(), theta_lp.numpy())
    model = ns["VI_SSM"](obs, obs_bin, tt, x0, x0_std, theta_dist, priors, dt, T, p, K, B, network_dims, B, F, fw,
                         learn_rate=1e-3, pre_train=False)
    model.build_flow()
    check_layout(st, layout, n)
    golden["lvb_mask_vals"] = np.asarray(model.mask_vals)
    golden["lvb_shift_vals"] = np.asarray(model.shift_vals)
    scale = float(B) / B
    loss, sde, obs_lp, prior_lp = model._ELBO()
    dev_obj = -(scale * (sde - model.lf_log_prob + obs_lp)).sum()
    g_theta = torch.autograd.grad(dev_obj, model.theta, retain_graph=True)[0]
    gv = ns["AdamaxOptimizer"](learning_rate=1e-3, beta1=0.95).compute_gradients(-loss)
    flat = torch.cat([gg.reshape(-1) for gg, _ in gv]).detach()
    golden.update({"lvb_hyper": np.array([p, K, B, F, fw, seed]), "lvb_dt": np.array(dt), "lvb_obs": obs, "lvb_obs_bin": obs_bin,
                   "lvb_time_till": tt, "lvb_eps": eps.numpy(), "lvb_theta": theta.numpy(), "lvb_theta_lp": theta_lp.numpy(),
                   "lvb_params_sha_f32": np.array(sha(params.numpy())), "lvb_sde": sde.detach().numpy(),
                   "lvb_obs_lp": obs_lp.detach().numpy(), "lvb_logq": model.lf_log_prob.detach().numpy(),
                   "lvb_prior": prior_lp.detach().numpy(), "lvb_lf_sample": model.lf_sample.detach().numpy(),
                   "lvb_elbo": loss.detach().numpy(), "lvb_grad_theta": g_theta.numpy()})
    grads_summary(flat, layout, "lvb_", golden)
    print("lv batch: sde", golden["lvb_sde"], "obs", golden["lvb_obs_lp"], "prior", golden["lvb_prior"], "global norm", float(flat.norm()))
    tf_shim.EVENT_REDUCTION = "literal"


def main():
    golden = {}
    if "--only-lvb" in sys.argv:
        golden = dict(np.load(os.path.join(HERE, "models_step_golden.npz")))
        lv_batch(golden)
        np.savez_compressed(os.path.join(HERE, "models_step_golden.npz"), **golden)
        return
    fhn(golden)
    sv(golden)
    lvr(golden)
    lv_fixed(golden)
    lv_batch(golden)
    path = os.path.join(HERE, "models_step_golden.npz")
    np.savez_compressed(path, **golden)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
