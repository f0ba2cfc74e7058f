"""Deterministic synthetic inputs shared by tests/golden/make_golden_models.py (which feeds them to the reference's own
unmodified scripts) and by the tests (which feed them to the oracle and to the CUDA path).

The reference ships no data for fitz_nag_NVP.py and lotka_volterra_partial_batch_fix_theta.py (their dat/*.txt are
missing from the repository) and only dat/SV.dat for SV_dense.py; the series below have the shapes those scripts
expect at their committed hyper-parameters.  numpy's legacy RandomState streams are stable across versions, so the
arrays are identical wherever they are rebuilt.
"""
import numpy as np


def fhn_inputs(target_dims: int = 1000000, obs_every: int = 10, seed: int = 11):
    """obs, obs_bin, time_till: [2, target_dims] float32 (fitz_nag_NVP.py:469-476; both components observed every
    `obs_every`-th step, observations held constant in between like AR_dat_gen.py:17-31 does)."""
    rs = np.random.RandomState(seed)
    t = np.arange(target_dims, dtype=np.float64)
    lat = np.stack([1.8 * np.sin(0.0123 * t) + 0.2 * rs.standard_normal(target_dims),
                    1.1 * np.cos(0.0123 * t) + 1.4 + 0.2 * rs.standard_normal(target_dims)])
    obs_bin = np.zeros((2, target_dims))
    obs_bin[:, ::obs_every] = 1.0
    hold = (np.arange(target_dims) // obs_every) * obs_every
    obs = lat[:, hold]
    nxt = np.minimum(hold + obs_every, target_dims - 1)
    time_till = np.tile(((nxt - np.arange(target_dims)) % obs_every) * 0.1, (2, 1))
    return obs.astype(np.float32), obs_bin.astype(np.float32), time_till.astype(np.float32)


def sv_prices(n: int = 1809, seed: int = 5):
    """A positive price-like series of the length of dat/SV.dat (SV_dense.py:404 keeps [300:])."""
    rs = np.random.RandomState(seed)
    vol = np.exp(-4.2 + 0.4 * np.sin(np.arange(n) / 90.0))
    return np.exp(np.cumsum(vol * rs.standard_normal(n)) + 0.3).astype(np.float32)


def lv_inputs(n_series: int = 4, batch_dims: int = 151, seed: int = 17):
    """obs, obs_bin, time_till: [2, n_series * batch_dims] float32 in the layout of
    dat/our_files/fix_theta/LV_*_dense_test.txt (lotka_volterra_partial_batch_fix_theta.py:649-651): series
    concatenated along time, -1 where unobserved (none in the dense files), populations around (91, 99)."""
    rs = np.random.RandomState(seed)
    n = n_series * batch_dims
    t = np.arange(n, dtype=np.float64)
    obs = np.stack([90.0 + 25.0 * np.sin(0.21 * t) + rs.standard_normal(n),
                    100.0 + 25.0 * np.cos(0.21 * t) + rs.standard_normal(n)])
    obs = np.maximum(obs, 2.0)
    obs_bin = np.ones((2, n))
    time_till = np.zeros((2, n))
    return obs.astype(np.float32), obs_bin.astype(np.float32), time_till.astype(np.float32)
