"""Host side of the mini-batch feed: padded base arrays and subsequence-index sampling.

The reference rebuilds a [p, L0, Cf] float64 `time_feats` tensor with ~14*p numpy slices every
iteration (AR.py:267-288).  Here the padded series are placed on the device ONCE and the kernels
gather windows themselves from the sampled start indices, so the per-iteration host work is the index
draw alone — which stays the reference's own legacy `np.random.choice` call so indices are bit-exact.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


def ar_base_arrays(obs, obs_bin, time_till, T: int, F: int, K: int, fw: int) -> List[np.ndarray]:
    """Base arrays of the AR model in the order `config.ar_config` expects.

    Equivalent to VI_SSM.__init__ (AR.py:135-150).  The reference keeps `feat_window` shifted copies of
    the observation series (`obs_pad_store[i] = zeros(P-i) ++ obs ++ zeros(i)`); one array padded with
    P zeros in front and fw-1 behind serves all of them: obs_pad_store[i][q] == obs_pad[q + i].
    """
    P = F * K + 1
    T = int(np.int32(T))
    obs = np.asarray(obs, dtype=np.float64)
    obs_bin = np.asarray(obs_bin, dtype=np.float64)
    time_till = np.asarray(time_till, dtype=np.float64)
    if obs.shape[0] != T or obs_bin.shape[0] != T or time_till.shape[0] != T:
        raise ValueError("series length must equal T (AR.py:135-150 assumes it)")
    obs_pad = np.concatenate((np.zeros(P), obs, np.zeros(max(fw - 1, 0))))
    bin_feats = np.concatenate((np.ones(P), np.zeros(T)))
    time_pad = np.concatenate((np.zeros(P), np.arange(T + 1, dtype=np.float64)))
    tt0 = float(time_till[0])
    time_till_pad = np.concatenate((np.arange(P + tt0, tt0, -1), time_till))
    obs_bin_pad = np.concatenate((np.zeros(P), obs_bin))
    return [obs_pad, bin_feats, time_pad, time_till_pad, obs_bin_pad]


def sample_indices(T: int, B: int, p: int, rng=None) -> np.ndarray:
    """p subsequence starts from arange(0, T, B); with replacement iff B*p >= T (AR.py:257-265).

    `rng=None` uses numpy's global legacy stream exactly like the reference."""
    rng = np.random if rng is None else rng
    return rng.choice(np.arange(0, T, B), size=p, replace=bool(B * p >= T)).astype(np.int64)


def partition_indices(idx: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Row sharding for N GPUs: contiguous slices of the injected index list."""
    per = (len(idx) + world - 1) // world
    return idx[rank * per:(rank + 1) * per]


def fhn_base_arrays(obs, obs_bin, time_till, dt: float, T: float, target_dims: int, F: int, K: int,
                    fw: int) -> List[np.ndarray]:
    """Base arrays of the FitzHugh-Nagumo model in the order `config.fhn_config` expects
    (fitz_nag_NVP.py:165,187-202).  obs, obs_bin, time_till: [2, target_dims]."""
    D = 2
    P2 = F * K + D
    obs_flat = np.reshape(np.asarray(obs, dtype=np.float64), -1, 'F')
    obs_pad = np.concatenate((np.zeros(P2), obs_flat, np.zeros(5 * max(fw - 1, 0))))
    bin_feats = np.concatenate((np.ones(P2), np.zeros(target_dims * D)))
    time_pad = np.concatenate((np.zeros(P2), np.repeat(np.arange(dt, T + dt, dt), D)))
    lead = np.reshape(np.repeat(np.arange(np.round(P2 * (dt / D), 1), -dt, -dt), D), (D, -1), 'F')
    tt = np.reshape(np.concatenate((lead, np.asarray(time_till, dtype=np.float64)), 1), -1, 'F')
    return [obs_pad, bin_feats, time_pad, tt, np.asarray(obs_bin, dtype=np.float64).reshape(-1)]


def rolling_var(x: np.ndarray, K: int) -> np.ndarray:
    """[np.var(x[i:i+K]) for i in range(len(x) - K)] (SV_dense.py:159-161) from prefix sums in float64.
    (A14 of SURVEY section 8a; the O(T*K) Python loop of the reference becomes O(T).)"""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0] - K
    if n <= 0:
        return np.zeros(0)
    # centre first: var is shift invariant and the prefix-sum form cancels catastrophically otherwise
    xc = x - x.mean()
    c1 = np.concatenate(([0.0], np.cumsum(xc)))
    c2 = np.concatenate(([0.0], np.cumsum(xc * xc)))
    s1 = c1[K:K + n] - c1[:n]
    s2 = c2[K:K + n] - c2[:n]
    return np.maximum(s2 / K - (s1 / K) ** 2, 0.0)


def sv_base_arrays(obs, dt: float, T: float, F: int, K: int, fw: int, exact_var: bool = True,
                   var_fn=None) -> List[np.ndarray]:
    """Base arrays of the SV model in the order `config.sv_config` expects (SV_dense.py:159-184).
    `exact_var=True` evaluates the rolling variances with the reference's own np.var loop (bit-exact);
    False uses the O(T) prefix-sum form (agrees to ~1e-12 relative)."""
    obs_in = np.asarray(obs)                 # the script keeps the series in float32 (NP_DTYPE, SV_dense.py:404) and
    obs = obs_in.astype(np.float64)          # np.var / the first differences run in THAT precision (pinned by sv_golden.npz)
    obs_pad = np.concatenate((np.zeros(F * K), obs, np.zeros(5 * max(fw - 1, 0))))
    time_pad = np.concatenate((np.zeros(F * K + 1), np.arange(0.1, T + dt, dt)))
    obs_diff = obs[1:] - obs[:-1]
    if var_fn is not None:       # e.g. the device kernel nma_rolling_var (engine.rolling_var): same float32 values
        diff_in = obs_in[1:] - obs_in[:-1]
        var_store = np.asarray(var_fn(obs_in, K))
        var_diff_store = np.asarray(var_fn(diff_in, K))
    elif exact_var:
        diff_in = obs_in[1:] - obs_in[:-1]
        var_store = np.array([np.var(obs_in[i:i + K]) for i in range(0, obs_in.shape[0] - K)])
        var_diff_store = np.array([np.var(diff_in[i:i + K]) for i in range(0, diff_in.shape[0] - K)])   # log below too
    else:
        var_store = rolling_var(obs, K)
        var_diff_store = rolling_var(obs_diff, K)
    var_pad = np.concatenate((np.zeros((F + 1) * K), var_store))
    var_diff_pad = np.concatenate((np.zeros((F + 1) * K), np.log(var_diff_store), np.zeros(1)))
    return [obs_pad, time_pad, var_pad, var_diff_pad]


def lvr_base_arrays(obs, obs_bin, time_till, dt: float, T: float, target_dims: int, F: int, K: int,
                    fw: int) -> List[np.ndarray]:
    """Base arrays of the learned-theta Lotka-Volterra model in the order `config.lvr_config` expects
    (lotka_volterra_partial.py:186-205).  obs, obs_bin, time_till: [2, target_dims]."""
    D = 2
    P2 = F * K + D
    obs_flat = np.reshape(np.asarray(obs, dtype=np.float64), -1, 'F')
    obs_pad = np.concatenate((np.zeros(P2), obs_flat, np.zeros(5 * max(fw - 1, 0))))
    bin_feats = np.concatenate((np.zeros(P2), np.ones(target_dims * D)))
    time_pad = np.concatenate((np.zeros(P2), np.repeat(np.arange(dt, T + dt, dt), D)))
    lead = np.reshape(np.repeat(np.arange(np.round(P2 * (dt / D), 1), 0., -dt), D), (D, -1), 'F')
    tt = np.reshape(np.concatenate((lead, np.asarray(time_till, dtype=np.float64)), 1), -1, 'F')
    return [obs_pad, bin_feats, time_pad, tt, np.asarray(obs_bin, dtype=np.float64).reshape(-1)]


def lv_base_arrays(obs, obs_bin, time_till, dt: float, T: float, target_dims: int, F: int, K: int, fw: int,
                   p_val: int = 1) -> List[np.ndarray]:
    """Base arrays of the Lotka-Volterra model in the order `config.lv_config` expects
    (lotka_volterra_partial_batch_fix_theta.py:186,203-222).  obs, obs_bin, time_till: [2, target_dims * p_val];
    unobserved entries of obs already replaced by 1 + softplus(-2) (ibid. :661-663)."""
    D = 2
    P2 = F * K + D
    obs_flat = np.reshape(np.asarray(obs, dtype=np.float64), -1, 'F')
    obs_pad = np.concatenate((np.zeros(P2), obs_flat, np.zeros(5 * max(fw - 1, 0))))
    bin_feats = np.concatenate((np.zeros(P2), np.ones(target_dims * D * p_val)))
    time_pad = np.concatenate((np.zeros(P2), np.repeat(np.arange(0, T + dt, dt), D * p_val)))
    lead = np.reshape(np.repeat(np.arange(np.round(P2 * (dt / D), 1), 0., -dt), D), (D, -1), 'F')
    tt = np.reshape(np.concatenate((lead, np.asarray(time_till, dtype=np.float64)), 1), -1, 'F')
    return [obs_pad, bin_feats, time_pad, tt, np.asarray(obs_bin, dtype=np.float64).reshape(-1)]


def sample_indices_lv(target_dims: int, B: int, p_val: int, rng=None) -> np.ndarray:
    """p_val subsequence starts from arange(0, target_dims * p_val, B), never with replacement
    (lotka_volterra_partial_batch_fix_theta.py:471,478-479)."""
    rng = np.random if rng is None else rng
    return rng.choice(np.arange(0, target_dims * p_val, B), size=p_val, replace=False).astype(np.int64)
