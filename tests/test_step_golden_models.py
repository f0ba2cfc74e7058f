"""FitzHugh-Nagumo and stochastic-volatility models: the oracle's flow / ELBO / gradient restatement against values
produced by the reference's OWN classes (tests/golden/models_step_golden.npz, made by
tests/golden/make_golden_step_models.py: the class sections of fitz_nag_NVP.py and SV_dense.py exec'd verbatim over
tests/golden/tf_shim.py).  See tests/test_step_golden.py for what such a fixture pins and what it does not.
CPU only; the GPU leg is tests/test_gpu_step_golden.py."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import nma_oracle as O
from viforssms_b200 import feed
from viforssms_b200.config import fhn_config, lv_config, lvr_config, param_layout, sv_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def GM():
    return np.load(os.path.join(ROOT, "tests", "golden", "models_step_golden.npz"))


def _params(cfg, layout, n, g, time_row, T):
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


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fhn_inputs(GM):
    p, K, B, F, fw, target_dims, seed = (int(v) for v in GM["fhn_hyper"])
    dt = float(GM["fhn_dt"])
    T = target_dims * dt
    cfg = fhn_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=target_dims, dt=dt)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    params = _params(cfg, layout, n, g, None, T)
    assert _sha(params.numpy()) == str(GM["fhn_params_sha_f32"])
    obs, obs_bin, tt, idx = GM["fhn_obs"], GM["fhn_obs_bin"], GM["fhn_time_till"], GM["fhn_idx"]
    pads = O.pad_series_fhn(obs, tt, np.array([2.0, 3.0]), dt, T, target_dims, F, K, fw)
    tf64, _, _, bin_feed = O.gather_feed_fhn(pads, obs_bin, idx, cfg.L0, B)
    f32 = lambda a: torch.from_numpy(np.asarray(a).astype(np.float32)).double()
    arrays = feed.fhn_base_arrays(obs, obs_bin, tt, dt, T, target_dims, F, K, fw)
    return (cfg, layout, n, params, torch.from_numpy(GM["fhn_eps"]), torch.from_numpy(GM["fhn_theta"]), idx, f32(tf64),
            {"bin_feed": f32(bin_feed)}, arrays)


def sv_inputs(GM):
    p, K, B, F, fw, N, seed = (int(v) for v in GM["sv_hyper"])
    dt, x0 = float(GM["sv_dt"]), float(GM["sv_x0"])
    T = float(N)
    cfg = sv_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=x0)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    params = _params(cfg, layout, n, g, fw, N)
    assert _sha(params.numpy()) == str(GM["sv_params_sha_f32"])
    obs, idx = GM["sv_obs"], GM["sv_idx"]
    pads = O.pad_series_sv(obs, x0, dt, T, N, F, K, fw)
    tf64, mask, shift, dim_one = O.gather_feed_sv(pads, idx, cfg.L0, B)
    f32 = lambda a: torch.from_numpy(np.asarray(a).astype(np.float32)).double()
    arrays = feed.sv_base_arrays(obs, dt, T, F, K, fw)
    return (cfg, layout, n, params, torch.from_numpy(GM["sv_eps"]), torch.from_numpy(GM["sv_theta"]), idx, f32(tf64),
            {"mask": f32(mask), "shift": f32(shift), "dim_one": f32(dim_one)}, arrays)


def lvr_inputs(GM):
    """lotka_volterra_partial.py on the dat/LV_*.txt series the reference ships."""
    p, K, B, F, fw, target_dims, seed = (int(v) for v in GM["lvr_hyper"])
    dt = float(GM["lvr_dt"])
    T = 50.0
    obs, obs_bin, tt = GM["lvr_obs"], GM["lvr_obs_bin"], GM["lvr_time_till"]      # the input series ride in the fixture
    x0 = np.array([100.0, 100.0])
    cfg = lvr_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=target_dims, dt=dt, x0=x0)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    params = _params(cfg, layout, n, g, None, T)
    for i in range(F):
        off, shape = layout[f"f{i}.feat0.w"]
        params[off:off + shape[0] * shape[1]] *= 0.02
    assert _sha(params.numpy()) == str(GM["lvr_params_sha_f32"])
    idx = GM["lvr_idx"]
    pads = O.pad_series_lvr(obs, tt, x0, dt, T, target_dims, F, K, fw)
    tf64, mask, shift, bin_feed = O.gather_feed_fhn(pads, obs_bin, idx, cfg.L0, B)
    f32 = lambda a: torch.from_numpy(np.asarray(a).astype(np.float32)).double()
    arrays = feed.lvr_base_arrays(obs, obs_bin, tt, dt, T, target_dims, F, K, fw)
    return (cfg, layout, n, params, torch.from_numpy(GM["lvr_eps"]), torch.from_numpy(GM["lvr_theta"]), idx, f32(tf64),
            {"mask": f32(mask), "shift": f32(shift), "bin_feed": f32(bin_feed)}, arrays)


def _close(got, want, rtol):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.abs(got - want).max() <= rtol * max(1.0, np.abs(want).max())


def check_grads(got, layout, GM, prefix, rtol, floor):
    """per-variable norms and leading entries; returns the worst relative error seen"""
    heads = GM[prefix + "grad_heads"]
    gn = float(GM[prefix + "global_norm"])
    pos, worst = 0, 0.0
    for nm, want_norm in zip([str(s) for s in GM[prefix + "var_names"]], GM[prefix + "grad_norms"]):
        off, shape = layout[nm]
        k = int(np.prod(shape))
        scale = max(want_norm, floor * gn)
        assert abs(np.linalg.norm(got[off:off + k]) - want_norm) <= rtol * scale, (prefix, nm)
        h = min(k, 16)
        hs = max(np.linalg.norm(heads[pos:pos + h]), floor * gn)
        err = np.linalg.norm(got[off:off + h] - heads[pos:pos + h])
        worst = max(worst, err / hs)
        assert err <= rtol * hs, (prefix, nm)
        pos += h
    assert abs(np.linalg.norm(got) - gn) <= rtol * gn
    return worst


def test_fhn_oracle_matches_the_reference_classes(GM):
    cfg, layout, n, params, eps, theta, idx, tf, extra, _ = fhn_inputs(GM)
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, extra=extra)
    t = ref["terms"].numpy()
    assert _close(t[:, 0], GM["fhn_sde"], 1e-10)              # fitz_nag_NVP.py:236-255
    assert _close(t[:, 1], GM["fhn_obs_lp"], 1e-10)           # :232-233
    assert _close(t[:, 2], GM["fhn_logq"], 1e-10)             # base log-density minus the coupling layers' log sigma
    assert _close(ref["lf"].numpy(), GM["fhn_lf_sample"], 1e-10)          # [p, 2, B+1] (:282-283)
    th = theta.double().numpy()
    prior = sum(-0.5 * (th[:, k] / 10.0) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(10.0) for k in range(5))
    assert _close(cfg.scale * (t[:, 0] - t[:, 2] + t[:, 1]) + prior, GM["fhn_elbo"], 1e-10)      # :265-268
    assert _close(ref["grad_theta"].numpy(), GM["fhn_grad_theta"], 1e-9)
    check_grads(ref["grad_params"].numpy(), layout, GM, "fhn_", 1e-9, 1e-12)
    pre = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, obj=2, extra=extra,
                           path_target=0.0)
    check_grads(pre["grad_params"].numpy(), layout, GM, "fhn_pre_", 1e-9, 1e-12)                # :288-289


def test_sv_oracle_matches_the_reference_classes(GM):
    cfg, layout, n, params, eps, theta, idx, tf, extra, _ = sv_inputs(GM)
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, extra=extra)
    t = ref["terms"].numpy()
    assert _close(t[:, 0], GM["sv_sde"], 1e-10)               # SV_dense.py:203-226
    assert _close(t[:, 2], GM["sv_logq"], 1e-10)
    assert _close(ref["lf"].numpy(), GM["sv_lf_sample"], 1e-10)           # [p, 2, B+1]: price, masked log-volatility
    th = theta.double().numpy()
    prior = sum(-0.5 * (th[:, k] / 10.0) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(10.0) for k in range(4))
    assert _close(cfg.scale * (t[:, 0] - t[:, 2]) + prior, GM["sv_elbo"], 1e-10)                 # :231-232
    assert _close(ref["grad_theta"].numpy(), GM["sv_grad_theta"], 1e-9)
    check_grads(ref["grad_params"].numpy(), layout, GM, "sv_", 1e-9, 1e-12)
    pre = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, obj=2, extra=extra,
                           path_target=-7.0)
    check_grads(pre["grad_params"].numpy(), layout, GM, "sv_pre_", 1e-9, 1e-12)                 # :251-252


def test_lvr_oracle_matches_the_reference_classes(GM):
    """lotka_volterra_partial.py (learned theta): transposed wide feature layer, coupling flow, softplus path with
    mask / shift, bivariate Euler-Maruyama density on the state differences with theta = exp(sample)."""
    cfg, layout, n, params, eps, theta, idx, tf, extra, arrays = lvr_inputs(GM)
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, extra=extra)
    t = ref["terms"].numpy()
    assert _close(t[:, 0], GM["lvr_sde"], 1e-10)              # lotka_volterra_partial.py:237-262
    assert _close(t[:, 1], GM["lvr_obs_lp"], 1e-10)           # :235
    assert _close(t[:, 2], GM["lvr_logq"], 1e-10)             # :291-293
    assert _close(ref["lf"].numpy(), GM["lvr_lf_sample"], 1e-10)
    th = theta.double().numpy()
    pri = [(np.log(4.428 / 10), 1e-4), (np.log(0.029 / 10), 1e-4), (np.log(2.957 / 10), 1e-4)]     # :476
    prior = sum(-0.5 * ((th[:, k] - m) / s) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(s) for k, (m, s) in enumerate(pri))
    assert _close(cfg.scale * (t[:, 0] - t[:, 2] + t[:, 1]) + prior, GM["lvr_elbo"], 1e-10)       # :270-271
    assert _close(ref["grad_theta"].numpy(), GM["lvr_grad_theta"], 1e-9)
    check_grads(ref["grad_params"].numpy(), layout, GM, "lvr_", 1e-9, 1e-12)
    pre = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, obj=2, extra=extra,
                           path_target=75.0)
    check_grads(pre["grad_params"].numpy(), layout, GM, "lvr_pre_", 1e-9, 1e-12)                 # :299-300
    # the host side builds the same base arrays as the oracle's padding (and so as the script)
    pads = O.pad_series_lvr(GM["lvr_obs"], GM["lvr_time_till"], np.array([100.0, 100.0]), cfg.dt, 50.0, 500, cfg.F, cfg.K, 10)
    assert np.array_equal(arrays[1], pads["bin_feats"]) and np.array_equal(arrays[2], pads["time_pad"])
    assert np.array_equal(arrays[3], pads["time_till"])
    assert np.array_equal(arrays[0][:pads["obs_pad_store"][0].shape[0]], pads["obs_pad_store"][0])


def lv_fixed_inputs(GM):
    p, K, B, F, fw, seed = (int(v) for v in GM["lvf_hyper"])
    dt = float(GM["lvf_dt"])
    N = p * B
    T = (N - 1) * dt
    cfg = lv_config(p=p, K=K, B=B, F=F, H=2, feat_window=fw, target_dims=B, dt=dt, x0=(11.0, 9.5))
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    params = _params(cfg, layout, n, g, None, T)
    for i in range(F):
        off, _ = layout[f"f{i}.head.b"]
        params[off] = 3.0
    assert _sha(params.numpy()) == str(GM["lvf_params_sha_f32"])
    obs, obs_bin, tt = GM["lvf_obs"], GM["lvf_obs_bin"], GM["lvf_time_till"]
    pads = O.pad_series_lv(obs, tt, np.array(cfg.x0), dt, T, N, 1, F, K, fw)
    idx = np.arange(p, dtype=np.int64) * B
    tf64, mask, shift, bin_feed = O.gather_feed_lv(pads, obs_bin, idx, cfg.L0, B)
    f32 = lambda a: torch.from_numpy(np.asarray(a).astype(np.float32)).double()
    theta = torch.from_numpy(GM["lvf_theta"]).float().repeat(p, 1)
    arrays = feed.lv_base_arrays(obs, obs_bin, tt, dt, T, N, F, K, fw, p_val=1)
    return (cfg, layout, n, params, torch.from_numpy(GM["lvf_eps"]), theta, idx, f32(tf64),
            {"mask": f32(mask), "shift": f32(shift), "bin_feed": f32(bin_feed)}, arrays)


def test_lv_fixed_theta_oracle_matches_the_reference_classes_under_the_per_state_reading(GM):
    """lotka_volterra_partial_batch_fix_theta.py: the script's classes run under both readings of how
    Softplus(event_ndims=2) reduces the log-determinant of the flattened [states, 2] matrix (see
    tests/golden/tf_shim.py, EVENT_REDUCTION).  The oracle - and the kernel - implement the per-state reading and agree
    with the script under it in every term; under the literal reading ONLY the transition term differs, by the
    log-determinants of the other states of the batch.  Which one TensorFlow 1.8 computes is the open question of
    DESIGN.md section 0."""
    cfg, layout, n, params, eps, theta, idx, tf, extra, _ = lv_fixed_inputs(GM)
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, extra=extra)
    t = ref["terms"].numpy()
    pre = "lvf_per_state_"
    assert _close(t[:, 0], GM[pre + "sde"], 1e-9)
    assert _close(t[:, 1], GM[pre + "obs_lp"], 1e-9)
    assert _close(t[:, 2], GM[pre + "logq"], 1e-9)
    assert _close(ref["lf"].numpy(), GM[pre + "lf_sample"], 1e-10)
    assert _close(cfg.scale * (t[:, 0] - t[:, 2] + t[:, 1]), GM[pre + "elbo"], 1e-9)             # no prior term: :334-337
    check_grads(ref["grad_params"].numpy(), layout, GM, pre, 1e-8, 1e-12)
    lit = "lvf_literal_"
    assert _close(t[:, 1], GM[lit + "obs_lp"], 1e-9) and _close(t[:, 2], GM[lit + "logq"], 1e-9)
    assert _close(ref["lf"].numpy(), GM[lit + "lf_sample"], 1e-10)
    assert not _close(t[:, 0], GM[lit + "sde"], 1e-3)


def lvb_inputs(GM):
    """Inputs of the lotka_volterra_partial_batch.py fixture (p_val = 3 windows tiling three concatenated 6-step series)."""
    from viforssms_b200.config import lvb_config
    p, K, B, F, fw, seed = (int(v) for v in GM["lvb_hyper"])
    dt = float(GM["lvb_dt"])
    N, T = p * B, (B - 1) * dt
    cfg = lvb_config(p=p, K=K, B=B, F=F, H=2, feat_window=fw, target_dims=B, dt=dt, x0=(11.0, 9.5))
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    obs, obs_bin, tt = GM["lvb_obs"], GM["lvb_obs_bin"], GM["lvb_time_till"]
    pads = O.pad_series_lv(obs, tt, np.array(cfg.x0), dt, T, B, p, F, K, fw)
    idx = np.arange(p, dtype=np.int64) * B
    tf64, mask, shift, bin_feed = O.gather_feed_lv(pads, obs_bin, idx, cfg.L0, B)
    params = _params(cfg, layout, n, g, None, T)
    for i in range(F):
        off, _ = layout[f"f{i}.head.b"]
        params[off] = 3.0
    assert _sha(params.numpy()) == str(GM["lvb_params_sha_f32"])
    f32 = lambda a: torch.from_numpy(np.asarray(a).astype(np.float32)).double()
    extra = {"mask": f32(mask), "shift": f32(shift), "bin_feed": f32(bin_feed)}
    arrays = feed.lv_base_arrays(obs, obs_bin, tt, dt, T, B, F, K, fw, p_val=p)
    eps = torch.from_numpy(GM["lvb_eps"])
    theta = torch.from_numpy(GM["lvb_theta"])
    return cfg, layout, n, params, eps, theta, idx, f32(tf64), extra, arrays


def test_lv_batch_oracle_matches_the_reference_classes(GM):
    """lotka_volterra_partial_batch.py: learned softplus-theta, plain bivariate transition density, p_val = 3 windows whose
    first p_val states are pinned - the oracle against the script's own classes run over the TF stand-in."""
    cfg, layout, n, params, eps, theta, idx, tf, extra, arrays = lvb_inputs(GM)
    assert np.array_equal(extra["mask"][0, :, :4].numpy(), [[0, 0, 0, 1]] * 2) and extra["mask"][1:].min() == 1   # :237-240
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, extra=extra)
    t = ref["terms"].numpy()
    assert _close(t[:, 0], GM["lvb_sde"], 1e-10) and _close(t[:, 1], GM["lvb_obs_lp"], 1e-10) and _close(t[:, 2], GM["lvb_logq"], 1e-10)
    assert _close(ref["lf"].numpy(), GM["lvb_lf_sample"], 1e-10)
    check_grads(ref["grad_params"].numpy(), layout, GM, "lvb_", 1e-9, 1e-12)
    assert _close(ref["grad_theta"].numpy(), GM["lvb_grad_theta"], 1e-9)
    priors = [(-1.0, np.sqrt(0.1)), (-6.0, np.sqrt(0.1)), (-1.0, np.sqrt(0.1)), (-2.0, np.sqrt(0.1))]
    prior = O.lvb_theta_prior(theta.double(), priors).numpy()
    assert _close(prior, GM["lvb_prior"], 1e-10)
    elbo = cfg.scale * (t[:, 0] - t[:, 2] + t[:, 1]) + prior - GM["lvb_theta_lp"]
    assert _close(elbo, GM["lvb_elbo"], 1e-10)
