"""CPU tests of the host side: the reference's CLI surface, time sharding + halo exchange over gloo
(world_size 2), and the index feed."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from oracle import nma_oracle as O
from viforssms_b200 import feed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _main_module():
    spec = importlib.util.spec_from_file_location("nma_main_cli", os.path.join(ROOT, "main.py"))
    m = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["main.py"]
    try:
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    return m


def test_cli_file_and_overrides(tmp_path):
    """main.py hyperparameters.txt with the reference's flags (main.py:14-22,112-128)."""
    m = _main_module()
    a = m.handle_opts([os.path.join(ROOT, "hyperparameters.txt")])
    h = m.resolve(a)
    assert (h["T"], h["impute"], h["x0"], h["obs_std"], h["p"]) == (5000, 1, 10.0, 1.0, 50)
    assert (h["kernel_len"], h["batch_dims"], h["no_flows"], h["feat_window"]) == (50, 50, 3, 10)
    assert h["network_dims"] == [50, 50, 50] and h["priors"] == [(0.0, 10.0)] * 3
    assert h["learn_rate"] == 1e-3 and h["grad_clip"] == 2.5e8
    assert np.array_equal(h["theta"], np.array([5.0, 0.5, 3.0]))
    a = m.handle_opts([os.path.join(ROOT, "hyperparameters.txt"), "-time", "700", "-i", "5", "-t", "1", "-theta", "0.8",
                       "-t", "0.5", "-xzero", "2", "-o", "0.3", "-kernel_len", "20", "-b", "25", "-feat_window", "4"])
    h = m.resolve(a)
    assert (h["T"], h["impute"], h["x0"], h["obs_std"], h["kernel_len"], h["batch_dims"], h["feat_window"]) == \
        (700, 5, 2.0, 0.3, 20, 25, 4)
    assert h["theta"].tolist() == ["1", "0.8", "0.5"]          # strings, like the reference; data_gen converts
    # -repair output is itself a valid hyper-parameter file
    f = tmp_path / "h.txt"
    f.write_text(m.DEFAULT_FILE)
    assert m.parseparams(str(f)) == m.parseparams(os.path.join(ROOT, "hyperparameters.txt"))
    with pytest.raises(SystemExit):
        m.resolve(m.handle_opts([str(tmp_path / "missing.txt")]))


def test_shard_bounds_partition_the_candidates():
    from viforssms_b200.trainer import shard_bounds
    for T, B, world in ((5000, 50, 2), (10 ** 8, 50, 8), (1300, 25, 3), (100, 50, 4)):
        cands = []
        prev = 0
        for r in range(world):
            t0, t1 = shard_bounds(T, B, world, r)
            assert t0 == prev and t0 % B == 0
            cands += list(range(t0, t1, B))
            prev = t1
        assert prev == T and cands == list(range(0, T, B))


def _window_from_local(arrays, cfg_chan, idx_local, L0):
    """What the device gather computes: time_feats[slot, c] = base[a_c][idx + slot + off_c] (zero outside)."""
    out = np.zeros((L0, len(cfg_chan)), dtype=np.float32)
    for c, (a, off) in enumerate(cfg_chan):
        arr = arrays[a]
        for s in range(L0):
            q = idx_local + s + off
            if 0 <= q < len(arr):
                out[s, c] = arr[q]
    return out


def _halo_worker(rank, world, port, T, B, F, K, fw, series, ret):
    import torch.distributed as dist
    from viforssms_b200.trainer import exchange_halos, local_base_arrays, shard_bounds
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        obs, obs_bin, tt = series
        t0, t1 = shard_bounds(T, B, world, rank)
        P = F * K + 1
        fill = torch.from_numpy(obs[t0:t1].copy())
        binary = torch.from_numpy(obs_bin[t0:t1].copy())
        till = torch.from_numpy(tt[t0:t1].copy())
        obs_ext, bin_ext, till_ext, tt0 = exchange_halos(fill, binary, till, P, fw, rank, world)
        arrays = [a.float().numpy() for a in local_base_arrays(obs_ext, bin_ext, till_ext, tt0, t0, t1, P, fw)]
        ret[rank] = (t0, t1, arrays)
        # gradient all-reduce is a SUM over ranks (the reference differentiates the sum over rows, AR.py:228-229)
        g = torch.full((5,), float(rank + 1))
        dist.all_reduce(g)
        assert g[0].item() == sum(range(1, world + 1))
    finally:
        dist.destroy_process_group()


def test_time_sharding_halo_exchange_gloo_world2():
    """Two ranks, each owning half of the series: after the halo exchange every window a rank can draw is
    bit-identical to the reference's window on the unsharded series (oracle gather, AR.py:267-288)."""
    import torch.multiprocessing as mp
    from viforssms_b200.config import ar_config
    T, B, F, K, fw = 1200, 20, 2, 15, 4
    rs = np.random.RandomState(5)
    obs = rs.normal(5.0, 2.0, T)
    obs_bin = (rs.uniform(size=T) < 0.6).astype(np.float64)
    tt = rs.randint(1, 5, size=T).astype(np.float64)
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_halo_worker, args=(world, port, T, B, F, K, fw, (obs, obs_bin, tt), ret), nprocs=world, join=True)
    cfg = ar_config(p=4, K=K, B=B, F=F, feat_window=fw, T=T)
    chan = list(zip(cfg.chan_array, cfg.chan_offset))
    pads = O.pad_series_ar(obs, obs_bin, tt, 10.0, T, F, K, fw)
    for rank in range(world):
        t0, t1, arrays = ret[rank]
        for idx in (t0, t0 + B, t1 - B):
            want, _, _ = O.gather_feed_ar(pads, np.array([idx]), cfg.L0, B)
            got = _window_from_local(arrays, chan, idx - t0, cfg.L0)
            assert np.array_equal(got, want[0].astype(np.float32)), (rank, idx)


def test_index_feeder_matches_reference_draw():
    from viforssms_b200.trainer import IndexFeeder
    cand = np.arange(0, 5000, 50)
    f = IndexFeeder(cand, rows=50, replace=False, seed=1, offset=0, pinned=False)
    try:
        got = [f.get().numpy().copy() for _ in range(3)]
    finally:
        f.close()
    rs = np.random.RandomState(1)
    for g in got:
        assert np.array_equal(g, rs.choice(cand, size=50, replace=False))
    np.random.seed(1)
    assert np.array_equal(feed.sample_indices(5000, 50, 50), got[0])


def test_fhn_feed_arrays_match_reference_restatement():
    """Product-side FHN base arrays + channel table reproduce the windows of fitz_nag_NVP.py:187-202,350-370."""
    from viforssms_b200.config import fhn_config
    rs = np.random.RandomState(0)
    N, dt = 240, 0.1
    T = N * dt
    for (K, B, F, fw) in ((8, 6, 3, 3), (20, 12, 3, 10)):
        cfg = fhn_config(p=4, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N, dt=dt)
        obs = rs.normal(size=(2, N))
        ob = (rs.uniform(size=(2, N)) < .3).astype(float)
        tt = rs.uniform(size=(2, N)).round(1)
        arrays = feed.fhn_base_arrays(obs, ob, tt, dt, T, N, F, K, fw)
        pads = O.pad_series_fhn(obs, tt, np.array([2., 3.]), dt, T, N, F, K, fw)
        idx = np.array([0, B, (N // B - 1) * B])
        tf, mask, shift, bf = O.gather_feed_fhn(pads, ob, idx, cfg.L0, cfg.B)
        assert tf.shape == (3, cfg.L0, fw + 3) and bf.shape == (3, 2, B)
        assert len(pads["time_till"]) == len(pads["time_pad"]) + 2          # the FHN quirk
        for r, i in enumerate(idx):
            for c in range(cfg.Cf):
                a = arrays[cfg.chan_array[c]]
                off = cfg.chan_offset[c]
                w = np.array([a[2 * i + s + off] if 0 <= 2 * i + s + off < len(a) else 0. for s in range(cfg.L0)])
                assert np.array_equal(w, tf[r, :, c]), (r, c)
            assert np.array_equal(arrays[4].reshape(2, N)[:, i:i + B], bf[r])


def test_sv_feed_arrays_and_rolling_variance():
    """Product-side SV base arrays reproduce the windows of SV_dense.py:159-184,305-328; the O(T) prefix-sum
    rolling variance (A14) agrees with the reference's np.var loop."""
    from viforssms_b200.config import sv_config
    rs = np.random.RandomState(21)
    N = 600
    obs = np.exp(rs.normal(0, 0.3, size=N + 1).cumsum() * 0.05 + 2.0)
    for (K, B, F, fw) in ((10, 7, 3, 2), (50, 52, 5, 5)):
        cfg = sv_config(p=3, K=K, B=B, F=F, H=3, feat_window=fw, target_dims=N)
        arrays = feed.sv_base_arrays(obs, 1.0, float(N), F, K, fw)
        fast = feed.sv_base_arrays(obs, 1.0, float(N), F, K, fw, exact_var=False)
        for a, b in zip(arrays, fast):
            assert np.allclose(a, b, rtol=1e-10, atol=1e-12)
        pads = O.pad_series_sv(obs, -8.5, 1.0, float(N), N, F, K, fw)
        idx = np.array([0, B, ((N - B - 1) // B) * B])
        tf, mask, shift, d1 = O.gather_feed_sv(pads, idx, cfg.L0, cfg.B)
        for r, i in enumerate(idx):
            for c in range(cfg.Cf):
                a = arrays[cfg.chan_array[c]]
                off = cfg.chan_offset[c]
                w = np.array([a[i + s + off] if 0 <= i + s + off < len(a) else 0. for s in range(cfg.L0)])
                assert np.array_equal(w, tf[r, :, c]), (r, c)
            assert np.array_equal(arrays[0][i + cfg.head_offset:i + cfg.head_offset + B + 1], d1[r])
    want = np.array([np.var(obs[i:i + 50]) for i in range(len(obs) - 50)])
    assert np.allclose(feed.rolling_var(obs, 50), want, rtol=1e-10)


def test_lvr_facade_constructor_and_base_arrays_on_the_host():
    """LVR_VI_SSM (lotka_volterra_partial.py:162-217): the constructor's derived sizes and the base arrays it hands the
    library, against the oracle's padding (which is the script's own, tests/test_step_golden_models.py)."""
    import torch
    from oracle import nma_oracle as O
    from viforssms_b200.theta_flow import ThetaFlow
    from viforssms_b200.vi_ssm_models import LVR_VI_SSM
    rs = np.random.RandomState(2)
    target_dims, dt, F, K, B, fw = 120, 0.1, 3, 6, 10, 4
    T = target_dims * dt
    obs = rs.uniform(1.0, 200.0, size=(2, target_dims))
    obs_bin = (rs.uniform(size=(2, target_dims)) < 0.1).astype(np.float64)
    tt = rs.uniform(0.1, 10.0, size=(2, target_dims)).round(1)
    x0 = np.array([100.0, 90.0])
    np.random.seed(1)
    flow = ThetaFlow(3, 4, base_loc=0.0, base_scale=1.0, activation="elu")
    m = LVR_VI_SSM(obs, obs_bin, tt, x0, flow, [(0.0, 1.0)] * 3, dt, T, 7, K, B, [50] * 5, target_dims, F, fw,
                   device=torch.device("cpu"))
    assert m.kernel_ext == K * F + 2 * B + 2 == m.cfg.L0
    assert m.cfg.dtheta == 3 and m.cfg.D == 2 and m.cfg.H == 3 and m.cfg.x0 == (100.0, 90.0)
    assert m.cfg.scale == target_dims / B
    arrays = m._base_arrays()
    pads = O.pad_series_lvr(obs, tt, x0, dt, T, target_dims, F, K, fw)
    assert np.array_equal(arrays[1], pads["bin_feats"])
    assert np.array_equal(arrays[2], pads["time_pad"])
    assert np.array_equal(arrays[3], pads["time_till"])
    for i, shifted in enumerate(pads["obs_pad_store"]):          # the look-ahead channels are ONE array read at offsets 5 i
        n = shifted.shape[0]
        assert np.array_equal(arrays[0][5 * i:5 * i + n], shifted)
    assert np.array_equal(arrays[4].reshape(2, -1), obs_bin)


def test_series_generators_write_the_layouts_the_scripts_read(tmp_path):
    """The reference ships no FHN / SV / fixed-theta LV series; the root scripts' generators
    produce them in the layouts the scripts load (fitz_nag_NVP.py:452-454, SV_dense.py:406,
    lotka_volterra_partial_batch_fix_theta.py:655-668)."""
    import numpy as np
    import fitz_nag_NVP as fhn
    import SV_dense as sv
    import lotka_volterra_partial_batch_fix_theta as lv
    obs, obs_bin, tt = fhn.generate(400, dat_dir=str(tmp_path))
    for name, want in (("fitz_nag_obs_partial.txt", obs), ("fitz_nag_obs_binary.txt", obs_bin), ("fitz_nag_time_till.txt", tt)):
        got = np.loadtxt(tmp_path / name)
        assert got.shape == (2, 400) and np.allclose(got, want)
    assert obs_bin[:, 9::10].all() and obs_bin.sum() == 2 * 40              # both components observed every 10th step
    assert np.allclose(tt[0, :10], 0.1 * np.arange(9, -1, -1))               # countdown to the next observation, in time units
    assert np.array_equal(obs[:, 0], obs[:, 9])                              # look-ahead fill: a step carries the NEXT observation
    assert np.isfinite(obs).all() and np.abs(obs).max() < 10.0
    s = sv.generate(500, path=str(tmp_path / "SV.dat"))
    assert np.allclose(np.loadtxt(tmp_path / "SV.dat"), s) and s.shape == (500,) and (s > 0).all()
    o = lv.generate(2, dat_dir=str(tmp_path))
    assert o.shape == (2, 2 * 151) and (o > 1.0).all()                       # observations are 1 + softplus(. - 1)
    assert np.loadtxt(tmp_path / "LV_obs_binary_dense_test.txt").all()
