"""CPU tests: the oracle's numpy half against the reference-generated golden vectors, the oracle's
torch half against fp64 autograd/gradcheck, and the product's host-side feed code against both."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import nma_oracle as O
from viforssms_b200 import feed
from viforssms_b200.config import ar_config, param_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _load_dat():
    d = os.path.join(ROOT, "dat")
    return (np.loadtxt(os.path.join(d, "AR_obs_partial.txt"), np.float32),
            np.loadtxt(os.path.join(d, "AR_obs_binary.txt"), np.float32),
            np.loadtxt(os.path.join(d, "AR_time_till.txt"), np.float32))


def test_data_gen_matches_reference_files(tmp_path, golden):
    """AR_dat_gen.data_gen under seed 1 reproduces the reference's committed dat/AR_*.txt byte for byte."""
    import importlib
    import AR_dat_gen
    importlib.reload(AR_dat_gen)            # re-seeds like a fresh `import AR_dat_gen` (AR_dat_gen.py:3)
    AR_dat_gen.data_gen(5000, 1, 10.0, np.array([5.0, 0.5, 3.0]), 1.0, dat_dir=str(tmp_path))
    want = dict(zip(golden["dat_names"].tolist(), golden["dat_sha256"].tolist()))
    for name, h in want.items():
        with open(tmp_path / "dat" / name, "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == h, name
        with open(os.path.join(ROOT, "dat", name), "rb") as f:      # the copies shipped for main.py
            assert hashlib.sha256(f.read()).hexdigest() == h, name


def test_data_gen_imputed(golden, tmp_path):
    import AR_dat_gen
    np.random.seed(7)
    fill, binary, till = AR_dat_gen.simulate(600, 5, 2.0, np.array([1.0, 0.8, 0.5]), 0.3)
    AR_dat_gen._write(str(tmp_path), fill, binary, till)
    for key, name in (("imp_obs", "AR_obs_partial.txt"), ("imp_obs_bin", "AR_obs_binary.txt"),
                      ("imp_time_till", "AR_time_till.txt")):
        got = np.loadtxt(tmp_path / "dat" / name)
        assert np.array_equal(got, golden[key]), name


def test_oracle_padding_matches_reference(golden):
    obs, obs_bin, tt = _load_dat()
    pads = O.pad_series_ar(obs, obs_bin, tt, 10.0, 5000, 3, 50, 10)
    assert np.array_equal(np.stack(pads["obs_pad_store"]), golden["obs_pad_store"])
    for k, g in (("time_pad", "time_pad"), ("bin_feats", "bin_feats"), ("obs_bin", "obs_bin_pad"),
                 ("time_till", "time_till_pad"), ("mask_vals", "mask_vals"), ("shift_vals", "shift_vals")):
        assert np.array_equal(pads[k], golden[g]), k


def test_oracle_index_draw_and_gather_match_reference(golden):
    obs, obs_bin, tt = _load_dat()
    pads = O.pad_series_ar(obs, obs_bin, tt, 10.0, 5000, 3, 50, 10)
    np.random.seed(1)
    for it in range(3):
        sel = O.sample_indices(np.int32(5000), 50, 50)
        assert np.array_equal(sel, golden["it%d_batch_select" % it])
        tf, mask, shift = O.gather_feed_ar(pads, sel, 201, 50)
        assert tf.dtype == np.float64
        assert _sha(tf) == str(golden["it%d_time_feats_sha256" % it])
        assert _sha(tf.astype(np.float32)) == str(golden["it%d_time_feats_f32_sha256" % it])
        assert np.array_equal(tf[:5], golden["it%d_time_feats_rows" % it])
        assert np.array_equal(mask, golden["it%d_mask" % it])
        assert np.array_equal(shift, golden["it%d_shift" % it])
    assert golden["it0_batch_select"][:10].tolist() == [4000, 4200, 1650, 4050, 4650, 850, 1800, 4100, 3450, 3250]


def test_oracle_gather_imputed_series(golden):
    pads = O.pad_series_ar(golden["imp_obs"], golden["imp_obs_bin"], golden["imp_time_till"], 2.0, 600, 2, 20, 4)
    tf, mask, shift = O.gather_feed_ar(pads, golden["imp_batch_select"], 2 * 20 + 25 + 1, 25)
    assert np.array_equal(tf, golden["imp_time_feats"])
    assert np.array_equal(mask, golden["imp_mask"])
    assert np.array_equal(shift, golden["imp_shift"])


def test_product_feed_matches_reference(golden):
    """viforssms_b200.feed builds ONE padded observation array; channel i must equal obs_pad_store[i]."""
    obs, obs_bin, tt = _load_dat()
    arrs = feed.ar_base_arrays(obs, obs_bin, tt, 5000, 3, 50, 10)
    P, fw, T = 151, 10, 5000
    for i in range(fw):
        assert np.array_equal(arrs[0][i:i + P + T], golden["obs_pad_store"][i])
    assert np.array_equal(arrs[1].astype(np.float32), golden["bin_feats"])
    assert np.array_equal(arrs[2], golden["time_pad"])
    assert np.array_equal(arrs[3], golden["time_till_pad"])
    assert np.array_equal(arrs[4], golden["obs_bin_pad"])
    np.random.seed(1)
    for it in range(3):
        assert np.array_equal(feed.sample_indices(5000, 50, 50), golden["it%d_batch_select" % it])
    # with replacement when B*p >= T (AR.py:257-260)
    np.random.seed(5)
    a = feed.sample_indices(1000, 50, 50)
    np.random.seed(5)
    b = np.random.choice(np.arange(0, 1000, 50), size=50, replace=True)
    assert np.array_equal(a, b)


def test_param_layout_counts():
    cfg = ar_config()
    layout, n = param_layout(cfg)
    assert n == 431706
    assert layout["f0.conv.w"][1] == (50, 51, 50)
    assert layout["f1.feat0.w"][0] == 143902


def _small_cfg(**kw):
    base = dict(p=3, K=4, B=5, F=2, H=1, feat_window=3, T=40)
    base.update(kw)
    return ar_config(**base)


def _rand_inputs(cfg, seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    layout, n = param_layout(cfg)
    params = O.glorot_init(layout, n, g, dtype)
    for name, (off, shape) in layout.items():           # non-zero biases so their gradients are exercised
        if name.endswith(".b"):
            params[off:off + int(np.prod(shape))] = 0.1 * torch.randn(int(np.prod(shape)), generator=g, dtype=dtype)
    eps = torch.randn(cfg.p, cfg.L0, generator=g, dtype=dtype)
    theta = torch.randn(cfg.p, cfg.dtheta, generator=g, dtype=dtype) * 0.3
    tf = torch.randn(cfg.p, cfg.L0, cfg.Cf, generator=g, dtype=dtype)
    tf[:, :, -1] = (tf[:, :, -1] > 0).to(dtype)
    return layout, params, eps, theta, tf


def test_oracle_shapes_ar_default():
    cfg = ar_config()
    layout, params, eps, theta, tf = _rand_inputs(cfg, dtype=torch.float32)
    x, logq = O.flow_forward(cfg, O.unpack_params(params, layout), eps, theta, tf)
    assert x.shape == (50, 51) and logq.shape == (50,)


def test_oracle_gradcheck_small():
    """fp64 autograd gradient vs central differences along random directions (full gradcheck over
    ~50k parameters would take minutes)."""
    cfg = _small_cfg()
    layout, params, eps, theta, tf = _rand_inputs(cfg)
    ref = O.step_reference(cfg, layout, params, eps, theta, tf)
    g = torch.Generator().manual_seed(11)

    def f(pv, th):
        return O.objective(cfg, 0, O.unpack_params(pv, layout), eps, th, tf)[0]
    for trial in range(4):
        dp = torch.randn(params.shape, generator=g, dtype=torch.float64)
        dth = torch.randn(theta.shape, generator=g, dtype=torch.float64)
        h = 1e-6
        num = (f(params + h * dp, theta + h * dth) - f(params - h * dp, theta - h * dth)) / (2 * h)
        ana = (ref["grad_params"] * dp).sum() + (ref["grad_theta"] * dth).sum()
        assert abs(num - ana) <= 1e-6 * max(1.0, abs(ana)), (trial, float(num), float(ana))


def test_oracle_fp32_tracks_fp64():
    cfg = _small_cfg(p=4, K=6, B=7, F=3)
    layout, params, eps, theta, tf = _rand_inputs(cfg)
    r64 = O.step_reference(cfg, layout, params, eps, theta, tf)
    r32 = O.step_reference(cfg, layout, params.float(), eps.float(), theta.float(), tf.float())
    assert torch.allclose(r32["terms"].double(), r64["terms"], rtol=1e-4, atol=1e-4)
    gn = r64["grad_params"].norm()
    assert (r32["grad_params"].double() - r64["grad_params"]).norm() / gn < 1e-4


def test_oracle_flow_is_locally_affine():
    """x_t depends on eps only through positions <= t + F*K window: finite receptive field."""
    cfg = _small_cfg()
    layout, params, eps, theta, tf = _rand_inputs(cfg)
    P = O.unpack_params(params, layout)
    x0, _ = O.flow_forward(cfg, P, eps, theta, tf)
    eps2 = eps.clone()
    eps2[:, -1] += 1.0            # perturb the last base-noise slot only
    x1, _ = O.flow_forward(cfg, P, eps2, theta, tf)
    assert torch.allclose(x0[:, :-1], x1[:, :-1]) and not torch.allclose(x0[:, -1], x1[:, -1])


def test_adamax_closed_form():
    """optimisers/adamax.py:51-57: no bias correction, eps inside the max."""
    w = torch.tensor([1.0, -2.0]); g = torch.tensor([0.5, -4.0])
    m = torch.zeros(2); v = torch.zeros(2)
    w1, m1, v1 = O.adamax_step(w, g, m, v, lr=0.1, beta1=0.9)
    assert torch.allclose(v1, 0.1 * g)
    assert torch.allclose(m1, g.abs())
    assert torch.allclose(w1, w - 0.1 * (0.1 * g) / g.abs())
    w2, m2, v2 = O.adamax_step(w1, torch.zeros(2), m1, v1, lr=0.1, beta1=0.9)
    assert torch.allclose(m2, 0.999 * m1 + 1e-8)
    # clip_by_global_norm (AR.py:230-232)
    w3, _, v3 = O.adamax_step(w, g, m, v, lr=0.1, beta1=0.9, clip=(1.0, float(g.norm())))
    assert torch.allclose(v3, 0.1 * g / g.norm())
