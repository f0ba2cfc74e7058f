"""GPU: the device gather (nma_gather through the C-ABI) against the feeds the reference's own FitzHugh-Nagumo,
stochastic-volatility and Lotka-Volterra scripts produced (tests/golden/*_golden.npz): bit-exact after the
float64 -> float32 feed_dict cast."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from viforssms_b200 import feed
from viforssms_b200.config import fhn_config, lv_config, sv_config

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden(name):
    return np.load(os.path.join(HERE, "golden", name), allow_pickle=False)


def _gather_check(cfg, arrays, g, tags, mask_shape):
    from viforssms_b200.engine import NMAEngine
    eng = NMAEngine(cfg)
    eng.set_series(arrays)
    for tag in tags:
        sel = g[tag + "_batch_select"].astype(np.int64)
        tf, mask, shift = eng.gather(sel)
        tf = tf.cpu().numpy()
        assert tf.dtype == np.float32 and list(tf.shape) == g[tag + "_time_feats_shape"].tolist()
        assert _sha(tf) == str(g[tag + "_time_feats_f32_sha256"]), tag
        n = g[tag + "_time_feats_rows"].shape[0]
        assert np.array_equal(tf[:n], g[tag + "_time_feats_rows"].astype(np.float32))
        assert np.array_equal(mask.cpu().numpy().reshape(mask_shape(len(sel))), g[tag + "_mask"].astype(np.float32))
        assert np.array_equal(shift.cpu().numpy().reshape(mask_shape(len(sel))), g[tag + "_shift"].astype(np.float32))
    eng.close()


def test_fhn_device_gather_matches_reference_feed():
    g = _golden("fhn_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs, obs_bin, tt = synth.fhn_inputs(N)
    dt, T = float(g["dt"]), float(g["T"])
    cfg = fhn_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt)
    cfg.x0 = tuple(float(v) for v in g["x0"])
    _gather_check(cfg, feed.fhn_base_arrays(obs, obs_bin, tt, dt, T, N, F, K, fw), g,
                  ("paths0", "paths1", "train0", "train1"), lambda n: (n, 2, B + 1))


def test_sv_device_gather_matches_reference_feed():
    g = _golden("sv_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs = synth.sv_prices()[300:]
    dt, T, x0 = float(g["dt"]), float(g["T"]), float(g["x0"])
    cfg = sv_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=x0)
    _gather_check(cfg, feed.sv_base_arrays(obs, dt, T, F, K, fw), g, ("paths0", "paths1", "train0", "train1"),
                  lambda n: (n, B + 1))


def test_lv_device_gather_matches_reference_feed():
    g = _golden("lv_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs, obs_bin, tt = synth.lv_inputs()
    obs = obs.copy()
    obs[obs == -1] = float(g["obs_not_observed"])
    dt, T = float(g["dt"]), float(g["T"])
    cfg = lv_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=g["x0_mean"])
    arrays = feed.lv_base_arrays(obs[:, :B], obs_bin[:, :B], tt[:, :B], dt, T, N, F, K, fw, p_val=p)
    _gather_check(cfg, arrays, g, ("paths0", "train0", "train1"), lambda n: (n, 2, B + 1))


def test_lv_batch_device_gather_matches_reference_feed():
    """lotka_volterra_partial_batch.py as committed (p_val = 3): time_feats, and the mask / shift that pin the first p_val
    states of the concatenated series, bit for bit what the script fed to its session."""
    from viforssms_b200.config import lvb_config
    g = _golden("lvb_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs, obs_bin, tt = synth.lv_inputs()
    obs = obs.copy()
    obs[obs == -1] = float(g["obs_not_observed"])
    dt, T = float(g["dt"]), float(g["T"])
    cfg = lvb_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=g["x0_mean"])
    sl = slice(0, p * B)
    arrays = feed.lv_base_arrays(obs[:, sl], obs_bin[:, sl], tt[:, sl], dt, T, N, F, K, fw, p_val=p)
    _gather_check(cfg, arrays, g, ("train0", "train1", "train2"), lambda n: (n, 2, B + 1))


def test_rolling_variance_kernel_is_bit_exact_with_numpy_float32():
    """A14 (SV_dense.py:159-170): nma_rolling_var reproduces np.var on the float32 series bit for bit - for the window of
    the script (50), below numpy's 8-wide unroll, and above its 128-element pairwise block - and yields the golden
    var_pad / var_diff_pad of the reference script."""
    from viforssms_b200.engine import rolling_var
    g = _golden("sv_golden.npz")
    obs = synth.sv_prices()[300:]
    dev = torch.device("cuda")

    def dev_var(x, K):
        return rolling_var(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev), K).cpu().numpy()
    for K in (50, 5, 8, 9, 129, 300):
        want = np.array([np.var(obs[i:i + K]) for i in range(obs.shape[0] - K)], dtype=np.float32)
        assert np.array_equal(dev_var(obs, K), want), K
    rs = np.random.RandomState(4)
    x = (rs.standard_normal(5000) * 37.0 - 11.0).astype(np.float32)
    want = np.array([np.var(x[i:i + 50]) for i in range(x.shape[0] - 50)], dtype=np.float32)
    assert np.array_equal(dev_var(x, 50), want)
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    arrays = feed.sv_base_arrays(obs, float(g["dt"]), float(g["T"]), F, K, fw, var_fn=dev_var)
    assert np.array_equal(arrays[2], g["var_pad"]) and np.array_equal(arrays[3], g["var_diff_pad"])
