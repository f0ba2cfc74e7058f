"""The numpy half of the FitzHugh-Nagumo, stochastic-volatility and Lotka-Volterra paths against what the reference's
own scripts fed to their TensorFlow session (tests/golden/{fhn,sv,lv}_golden.npz, made by
tests/golden/make_golden_models.py from the unmodified fitz_nag_NVP.py / SV_dense.py /
lotka_volterra_partial_batch_fix_theta.py): series padding, subsequence sampling and the window gather.

CPU tests check (1) the oracle restatement and (2) the product's host side - the base arrays of
`viforssms_b200.feed` read through the channel tables of `viforssms_b200.config`, i.e. the exact addressing the
device gather performs; the GPU tests run the device gather itself (nma_gather, through the C-ABI)."""
import hashlib
import os
import sys

import numpy as np
import pytest

from oracle import nma_oracle as O
from viforssms_b200 import feed
from viforssms_b200.config import fhn_config, lv_config, sv_config

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import synth  # noqa: E402


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden(name):
    return np.load(os.path.join(HERE, "golden", name), allow_pickle=False)


def table_gather(cfg, arrays, idx):
    """time_feats[r, j, c] = base[chan_array[c]][D*idx[r] + j + chan_offset[c]] (zero outside the array): what
    nma_gather computes (viforssms_b200/csrc/nma_fwd.cu, k_gather), restated in numpy on the float64 base arrays."""
    out = np.zeros((len(idx), cfg.L0, cfg.Cf))
    j = np.arange(cfg.L0)
    for c in range(cfg.Cf):
        base = np.asarray(arrays[cfg.chan_array[c]], dtype=np.float64)
        q = cfg.D * np.asarray(idx)[:, None] + j[None, :] + cfg.chan_offset[c]
        ok = (q >= 0) & (q < base.shape[0])
        out[:, :, c] = np.where(ok, base[np.clip(q, 0, base.shape[0] - 1)], 0.0)
    return out


def _check_feed(g, tag, tf, mask, shift, extra):
    assert list(tf.shape) == g[tag + "_time_feats_shape"].tolist()
    assert tf.dtype == np.float64
    assert _sha(tf) == str(g[tag + "_time_feats_sha256"]), tag
    assert _sha(tf.astype(np.float32)) == str(g[tag + "_time_feats_f32_sha256"]), tag
    n = g[tag + "_time_feats_rows"].shape[0]
    assert np.array_equal(tf[:n], g[tag + "_time_feats_rows"])
    assert np.array_equal(mask, g[tag + "_mask"]) and np.array_equal(shift, g[tag + "_shift"])
    for name, val in extra.items():
        assert np.array_equal(val, g[tag + "_" + name]), (tag, name)


# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def fhn_case():
    g = _golden("fhn_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs, obs_bin, tt = synth.fhn_inputs(N)
    return g, (p, K, B, F, N, fw), (obs, obs_bin, tt), float(g["dt"]), float(g["T"]), g["x0"]


def test_fhn_hyperparameters_are_the_scripts(fhn_case):
    g, hyper, _, dt, T, x0 = fhn_case
    assert hyper == (50, 20, 50, 3, 1000000, 10) and dt == 0.1 and T == 100000.0 and x0.tolist() == [2.0, 3.0]


def test_fhn_oracle_feed_matches_reference(fhn_case):
    g, (p, K, B, F, N, fw), (obs, obs_bin, tt), dt, T, x0 = fhn_case
    pads = O.pad_series_fhn(obs, tt, x0, dt, T, N, F, K, fw)
    L0 = F * K + 2 * B + 2
    np.random.seed(101)
    for tag in ("paths0", "paths1", "train0", "train1"):
        sel = g[tag + "_batch_select"]
        if tag.startswith("train"):     # the draw itself: same legacy stream, same call (fitz_nag_NVP.py:347-348)
            drawn = np.random.choice(np.arange(0, N, B), size=p, replace=bool(B * p >= N))
            assert np.array_equal(drawn, sel)
        tf, mask, shift, bin_feed = O.gather_feed_fhn(pads, obs_bin, sel, L0, B)
        _check_feed(g, tag, tf, mask, shift, {"bin_feed": bin_feed})


def test_fhn_product_feed_matches_reference(fhn_case):
    g, (p, K, B, F, N, fw), (obs, obs_bin, tt), dt, T, x0 = fhn_case
    cfg = fhn_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt)
    arrays = feed.fhn_base_arrays(obs, obs_bin, tt, dt, T, N, F, K, fw)
    for tag in ("paths1", "train0", "train1"):
        sel = g[tag + "_batch_select"]
        tf = table_gather(cfg, arrays, sel)
        assert _sha(tf) == str(g[tag + "_time_feats_sha256"]), tag
        # the observation indicator the ELBO reads: array 4 = obs_bin [2, N] row-major at idx + t (fitz_nag_NVP.py:369-370)
        ob = np.asarray(arrays[cfg.bin_array]).reshape(2, N)
        bin_feed = np.stack([ob[:, i:i + B] for i in sel])
        assert np.array_equal(bin_feed, g[tag + "_bin_feed"])


# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def sv_case():
    g = _golden("sv_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs = synth.sv_prices()[300:]
    return g, (p, K, B, F, N, fw), obs, float(g["dt"]), float(g["T"]), float(g["x0"])


def test_sv_oracle_feed_matches_reference(sv_case):
    g, (p, K, B, F, N, fw), obs, dt, T, x0 = sv_case
    assert (p, K, B, F, N, fw) == (200, 50, 52, 5, 1508, 5) and T == 1508 and x0 == -8.5
    pads = O.pad_series_sv(obs, x0, dt, T, N, F, K, fw)
    assert np.array_equal(pads["var_pad"], g["var_pad"]) and np.array_equal(pads["var_diff_pad"], g["var_diff_pad"])
    L0 = F * K + B + 1
    np.random.seed(202)
    for tag in ("paths0", "paths1", "train0", "train1"):
        sel = g[tag + "_batch_select"]
        if tag.startswith("train"):
            drawn = np.random.choice(np.arange(0, N, B), size=p, replace=bool(B * p >= N))
            assert np.array_equal(drawn, sel)
        tf, mask, shift, dim_one = O.gather_feed_sv(pads, sel, L0, B)
        _check_feed(g, tag, tf, mask, shift, {"dim_one": dim_one})


def test_sv_product_feed_matches_reference(sv_case):
    g, (p, K, B, F, N, fw), obs, dt, T, x0 = sv_case
    cfg = sv_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=x0)
    arrays = feed.sv_base_arrays(obs, dt, T, F, K, fw)          # exact_var: the reference's own np.var loop
    assert np.array_equal(arrays[2], g["var_pad"]) and np.array_equal(arrays[3], g["var_diff_pad"])
    for tag in ("paths1", "train0", "train1"):
        sel = g[tag + "_batch_select"]
        tf = table_gather(cfg, arrays, sel)
        assert _sha(tf) == str(g[tag + "_time_feats_sha256"]), tag
        base = np.asarray(arrays[cfg.obs_array])
        dim_one = np.stack([base[cfg.head_offset + i: cfg.head_offset + i + B + 1] for i in sel])
        assert np.array_equal(dim_one, g[tag + "_dim_one"])
    # the O(T) prefix-sum rolling variance (A14, float64) agrees with the reference's O(T*K) loop to the rounding of
    # the reference's own float32 np.var
    fast = feed.sv_base_arrays(obs, dt, T, F, K, fw, exact_var=False)
    assert np.allclose(fast[2], g["var_pad"], rtol=2e-5, atol=1e-9)
    assert np.allclose(fast[3], g["var_diff_pad"], rtol=0, atol=2e-5)


# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def lv_case():
    g = _golden("lv_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs, obs_bin, tt = synth.lv_inputs()
    obs = obs.copy()
    obs[obs == -1] = float(g["obs_not_observed"])
    sl = slice(0, B)                        # series 0 of the concatenated file (idx = 0 in the script's loop)
    return g, (p, K, B, F, N, fw), (obs[:, sl], obs_bin[:, sl], tt[:, sl]), float(g["dt"]), float(g["T"]), g["x0_mean"]


def test_lv_oracle_feed_matches_reference(lv_case):
    g, (p, K, B, F, N, fw), (obs, obs_bin, tt), dt, T, x0 = lv_case
    assert (p, K, B, F, N, fw) == (1, 20, 151, 3, 151, 10) and dt == 0.2 and T == 30
    assert np.allclose(g["priors"], np.log1p(np.exp([-1.0, -6.0, -1.0, -2.0])))
    pads = O.pad_series_lv(obs, tt, x0, dt, T, N, p, F, K, fw)
    L0 = F * K + 2 * B + 2
    for tag in ("paths0", "train0", "train1"):
        sel = g[tag + "_batch_select"]
        assert set(sel.tolist()) <= set(np.arange(0, N * p, B).tolist())
        tf, mask, shift, bin_feed = O.gather_feed_lv(pads, obs_bin, sel, L0, B)
        _check_feed(g, tag, tf, mask, shift, {"bin_feed": bin_feed})
    assert np.array_equal(O.sample_indices_lv(N, B, p), [0])


def test_lv_product_feed_matches_reference(lv_case):
    g, (p, K, B, F, N, fw), (obs, obs_bin, tt), dt, T, x0 = lv_case
    cfg = lv_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=x0)
    assert cfg.L0 == 364
    arrays = feed.lv_base_arrays(obs, obs_bin, tt, dt, T, N, F, K, fw, p_val=p)
    tf = table_gather(cfg, arrays, g["train0_batch_select"])
    assert _sha(tf) == str(g["train0_time_feats_sha256"])
    assert np.array_equal(tf, g["time_feats_full"])
    assert np.array_equal(feed.sample_indices_lv(N, B, p), [0])


# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def lvb_case():
    """lotka_volterra_partial_batch.py as committed: p_val = 3 windows over the first three concatenated series."""
    g = _golden("lvb_golden.npz")
    p, K, B, F, N, fw = (int(v) for v in g["hyper"])
    obs, obs_bin, tt = synth.lv_inputs()
    obs = obs.copy()
    obs[obs == -1] = float(g["obs_not_observed"])
    sl = slice(0, p * B)                    # obs[:, :p_val * batch_dims] (:717-719)
    return g, (p, K, B, F, N, fw), (obs[:, sl], obs_bin[:, sl], tt[:, sl]), float(g["dt"]), float(g["T"]), g["x0_mean"]


def test_lv_batch_oracle_feed_matches_reference(lvb_case):
    g, (p, K, B, F, N, fw), (obs, obs_bin, tt), dt, T, x0 = lvb_case
    assert (p, K, B, F, N, fw) == (3, 20, 151, 3, 151, 10) and dt == 0.2 and T == 30
    pads = O.pad_series_lv(obs, tt, x0, dt, T, N, p, F, K, fw)
    L0 = F * K + 2 * B + 2
    for tag in ("train0", "train1", "train2"):
        sel = g[tag + "_batch_select"]
        assert sorted(sel.tolist()) == [0, B, 2 * B]              # without replacement: always all p_val windows
        tf, mask, shift, bin_feed = O.gather_feed_lv(pads, obs_bin, sel, L0, B)
        _check_feed(g, tag, tf, mask, shift, {"bin_feed": bin_feed})
        # the FIRST p_val states of the concatenated series are pinned (mask_vals, :237-240): only the window starting at 0
        r0 = sel.tolist().index(0)
        assert mask[r0, :, :p].max() == 0 and mask[r0, :, p:].min() == 1 and np.delete(mask, r0, 0).min() == 1


def test_lv_batch_product_feed_matches_reference(lvb_case):
    from viforssms_b200.config import lvb_config
    g, (p, K, B, F, N, fw), (obs, obs_bin, tt), dt, T, x0 = lvb_case
    cfg = lvb_config(p=p, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt, x0=x0)
    assert cfg.L0 == 364 and cfg.n_pinned == p
    arrays = feed.lv_base_arrays(obs, obs_bin, tt, dt, T, N, F, K, fw, p_val=p)
    for tag in ("train0", "train1", "train2"):
        tf = table_gather(cfg, arrays, g[tag + "_batch_select"])
        assert _sha(tf) == str(g[tag + "_time_feats_sha256"])
    np.random.seed(5)
    assert sorted(feed.sample_indices_lv(N, B, p).tolist()) == [0, B, 2 * B]
