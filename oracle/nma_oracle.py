"""CPU oracle for the NMA ELBO step  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  The product path
(`viforssms_b200/`) never does; it fails loudly when the CUDA library is missing.

What it restates.  The reference (mehrnazmo/VIforSSMs) is TensorFlow-1.8 Python;
TensorFlow cannot be installed here (no wheel for CPython 3.12, no network), so
the arithmetic that TF's own kernels would run is restated with plain
numpy / torch-CPU ops, following the reference scripts line by line:

  pad_series_ar        AR.py:135-150           series padding (numpy, float64)
  sample_indices       AR.py:257-265           np.random.choice index draw
  gather_feed_ar       AR.py:267-288           window gather -> time_feats/mask/shift
  pad_series_fhn/_sv/_lv, gather_feed_fhn/_sv/_lv  the same for fitz_nag_NVP.py:187-202,350-370, SV_dense.py:159-184,
                       305-328 and lotka_volterra_partial_batch_fix_theta.py:203-222,478-503
  flow_forward         AR.py:24-35,44-110      base dist, IAF layer, Flow_Stack
                       fitz_nag_NVP.py:56-156  stride-2 head, interleave, Permute
                       SV_dense.py:37-89       delta-augmented features
  elbo_terms           AR.py:168-187           (+ fitz_nag_NVP.py:232-266, SV_dense.py:203-234)
  adamax_step          optimisers/adamax.py:42-58 + AR.py:226-234 (global-norm clip)

PARITY STATUS.  The numpy half (padding, index draw, gather, data generation) is
PINNED: tests/golden/{ar,fhn,sv,lv}_golden.npz hold what the reference's own
unmodified code fed to its session (tests/golden/make_golden.py and
make_golden_models.py run AR.py / fitz_nag_NVP.py / SV_dense.py /
lotka_volterra_partial_batch_fix_theta.py under a stub `tensorflow`).  The torch half:
  * AR model (flow, ELBO terms, gradients of -ELBO and of the pre-training objective, clip + Adamax): PINNED TO THE
    REFERENCE'S OWN CLASSES.  tests/golden/make_golden_step.py imports AR.py and optimisers/adamax.py unmodified and
    executes init_dist, IAF._create_flow, Flow_Stack, VI_SSM._ELBO / build_flow and AdamaxOptimizer over
    tests/golden/tf_shim.py, a torch float64 stand-in for the ~30 TensorFlow-1.8 LIBRARY ops they call; this
    restatement agrees with what they return to 1e-10 (terms, path) and 1e-9 (every gradient entry)
    (tests/test_step_golden.py; the CUDA path is held to 1e-4 against the same file, tests/test_gpu_step_golden.py).
    Pinned: the composition - slices, terms, signs, scales, variable creation order, slot arithmetic.  Not pinned:
    TensorFlow's own op kernels (the shim restates their documented behaviour) - the real TF cannot run here.
  * FHN and SV models: the same, from the class sections of fitz_nag_NVP.py and SV_dense.py exec'd verbatim
    (tests/golden/make_golden_step_models.py -> models_step_golden.npz; tests/test_step_golden_models.py): terms, path,
    ELBO, gradients of -ELBO and of the scripts' pre-training objective, 1e-10 / 1e-9.
  * Both Lotka-Volterra scripts likewise (lvr_* = lotka_volterra_partial.py, lvf_* = ..._batch_fix_theta.py); the
    fixed-theta script under BOTH readings of how Softplus(event_ndims=2) reduces the log-determinant of a flattened
    [states, 2] matrix - this restatement implements the per-state reading and equals the script's classes under it.
  * Still "parity unpinned": the theta posterior (tf.contrib masked-autoregressive-flow chains; restated on the host in
    viforssms_b200/theta_flow.py, executed from the reference nowhere) and which of the two event_ndims readings
    TensorFlow 1.8 really computes.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as Fnn

LOG2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------
# numpy half: padding, sampling, gather (pinned by tests/golden/ar_golden.npz)
# ----------------------------------------------------------------------------

def pad_series_ar(obs, obs_bin, time_till, x0, T, F, K, fw) -> Dict[str, object]:
    """AR.py:135-150.  Returns the same seven padded objects the reference keeps on `self`."""
    P = F * K + 1
    T = int(np.int32(T))
    store = []
    for i in range(fw):
        store.append(np.concatenate((np.zeros(P - i), obs, np.zeros(i)), axis=0))
    return {
        "obs_pad_store": store,
        "time_pad": np.concatenate((np.zeros(P), np.arange(T + 1)), axis=0),
        "bin_feats": np.float32(np.concatenate((np.ones(P), np.zeros(T)), axis=0)),
        "obs_bin": np.concatenate((np.zeros(P), obs_bin), axis=0),
        "mask_vals": np.concatenate((np.zeros((1, 1)), np.ones((1, T))), axis=1),
        "shift_vals": np.concatenate((np.array([[x0]]), np.zeros((1, T))), axis=1),
        "time_till": np.concatenate((np.arange(P + time_till[0], time_till[0], -1), time_till), axis=0),
    }


def sample_indices(T, B, p, rng=np.random) -> np.ndarray:
    """AR.py:257-265: p subsequence starts drawn from arange(0,T,B), with replacement iff B*p >= T."""
    cand = np.arange(0, T, B)
    return rng.choice(cand, size=p, replace=bool(B * p >= T))


def gather_feed_ar(pads, batch_select, L0, B) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """AR.py:267-288: (time_feats [p,L0,fw+4], mask [p,B+1], shift [p,B+1]), float64 like the feed."""
    def win(arr):
        return np.stack([arr[i:i + L0] for i in batch_select], axis=0)[:, :, None]
    chans = [win(a) for a in pads["obs_pad_store"]]
    chans += [win(pads["bin_feats"]), win(pads["time_pad"]), win(pads["time_till"]), win(pads["obs_bin"])]
    time_feats = np.concatenate(chans, axis=2)
    mask = np.stack([pads["mask_vals"][0, i:i + B + 1] for i in batch_select], axis=0)
    shift = np.stack([pads["shift_vals"][0, i:i + B + 1] for i in batch_select], axis=0)
    return time_feats, mask, shift


def pad_series_fhn(obs, time_till, x0, dt, T, target_dims, F, K, fw) -> Dict[str, object]:
    """fitz_nag_NVP.py:165,187-202 (flow_dims = 2; the two components interleave on the time axis)."""
    D = 2
    obs_flatten = np.reshape(obs, -1, 'F')
    store = []
    for i in range(0, fw * 5, 5):
        store.append(np.concatenate((np.zeros(F * K + D - i), obs_flatten, np.zeros(i)), axis=0))
    time_pad = np.concatenate((np.zeros(F * K + D), np.repeat(np.arange(dt, T + dt, dt), D)), axis=0)
    time_till_pad = np.reshape(np.repeat(np.arange(np.round((F * K + D) * (dt / D), 1), -dt, -dt), D), (D, -1), 'F')
    return {
        "obs_pad_store": store,
        "time_pad": time_pad,
        "time_till": np.reshape(np.concatenate((time_till_pad, time_till), 1), -1, 'F'),
        "bin_feats": np.float32(np.concatenate((np.ones(F * K + D), np.zeros(target_dims * D)), axis=0)),
        "mask_vals": np.concatenate((np.zeros((2, 1)), np.ones((D, target_dims))), axis=1),
        "shift_vals": np.concatenate((np.expand_dims(x0, 1), np.zeros((D, target_dims))), axis=1),
    }


def gather_feed_fhn(pads, obs_bin, batch_select, L0, B):
    """fitz_nag_NVP.py:350-370: (time_feats [p,L0,fw+3], mask, shift [p,2,B+1], bin_feed [p,2,B])."""
    def win(arr):
        return np.stack([arr[i:i + L0] for i in 2 * batch_select], axis=0)[:, :, None]
    chans = [win(a) for a in pads["obs_pad_store"]]
    chans += [win(pads["bin_feats"]), win(pads["time_pad"]), win(pads["time_till"])]
    time_feats = np.concatenate(chans, axis=2)
    mask = np.stack([pads["mask_vals"][:, i:i + B + 1] for i in batch_select], axis=0)
    shift = np.stack([pads["shift_vals"][:, i:i + B + 1] for i in batch_select], axis=0)
    bin_feed = np.stack([obs_bin[:, i:i + B] for i in batch_select], axis=0)
    return time_feats, mask, shift, bin_feed


def pad_series_sv(obs, x0, dt, T, target_dims, F, K, fw) -> Dict[str, object]:
    """SV_dense.py:159-184."""
    var_store = []
    for i in range(0, obs.shape[0] - K):
        var_store.append(np.var(obs[i:i + K]))
    var_pad = np.concatenate((np.zeros((F + 1) * K), var_store), axis=0)
    var_diff_store = []
    obs_diff = obs[1:] - obs[:-1]
    for i in range(0, obs_diff.shape[0] - K):
        var_diff_store.append(np.var(obs_diff[i:i + K]))
    var_diff_pad = np.concatenate((np.zeros((F + 1) * K), np.log(var_diff_store), np.zeros(1)), axis=0)
    store = []
    for i in range(0, fw * 5, 5):
        store.append(np.concatenate((np.zeros(F * K - i), obs, np.zeros(i)), axis=0))
    return {
        "obs": obs, "obs_pad_store": store, "var_pad": var_pad, "var_diff_pad": var_diff_pad,
        "time_pad": np.concatenate((np.zeros(F * K + 1), np.arange(0.1, T + dt, dt)), axis=0),
        "mask_vals": np.concatenate((np.zeros((1, 1)), np.ones((1, target_dims))), axis=1),
        "shift_vals": np.concatenate((np.array([[x0]]), np.zeros((1, target_dims))), axis=1),
    }


def gather_feed_sv(pads, batch_select, L0, B):
    """SV_dense.py:305-328: (time_feats [p,L0,fw+3], mask, shift, dim_one [p,B+1])."""
    def win(arr):
        return np.stack([arr[i:i + L0] for i in batch_select], axis=0)[:, :, None]
    chans = [win(a) for a in pads["obs_pad_store"]]
    chans += [win(pads["time_pad"]), win(pads["var_pad"]), win(pads["var_diff_pad"])]
    time_feats = np.concatenate(chans, axis=2)
    mask = np.stack([pads["mask_vals"][0, i:i + B + 1] for i in batch_select], axis=0)
    shift = np.stack([pads["shift_vals"][0, i:i + B + 1] for i in batch_select], axis=0)
    dim_one = np.stack([pads["obs"][i:i + B + 1] for i in batch_select], axis=0)
    return time_feats, mask, shift, dim_one


def pad_series_lv(obs, time_till, x0_mean, dt, T, target_dims, p_val, F, K, fw) -> Dict[str, object]:
    """lotka_volterra_partial_batch_fix_theta.py:186,203-222 (flow_dims = 2).  Unlike the FHN script the time channel
    starts at 0 and is repeated 2*p_val times, the lead of time_till stops before 0, and bin_feats is 0 on the pad and
    1 on the series."""
    D = 2
    obs_flatten = np.reshape(obs, -1, 'F')
    store = []
    for i in range(0, fw * 5, 5):
        store.append(np.concatenate((np.zeros(F * K + D - i), obs_flatten, np.zeros(i)), axis=0))
    time_pad = np.concatenate((np.zeros(F * K + D), np.repeat(np.arange(0, T + dt, dt), D * p_val)), axis=0)
    time_till_pad = np.reshape(np.repeat(np.arange(np.round((F * K + D) * (dt / D), 1), 0., -dt), D), (D, -1), 'F')
    return {
        "obs_pad_store": store,
        "time_pad": time_pad,
        "time_till": np.reshape(np.concatenate((time_till_pad, time_till), 1), -1, 'F'),
        "bin_feats": np.float32(np.concatenate((np.zeros(F * K + D), np.ones(target_dims * D * p_val)), axis=0)),
        "mask_vals": np.concatenate((np.zeros((2, p_val)), np.ones((D, target_dims * p_val))), axis=1),
        "shift_vals": np.concatenate((np.repeat(np.expand_dims(x0_mean, 1), p_val, 1),
                                      np.zeros((D, target_dims * p_val))), axis=1),
    }


def sample_indices_lv(target_dims, B, p_val, rng=np.random) -> np.ndarray:
    """lotka_volterra_partial_batch_fix_theta.py:478-479: always without replacement, candidates over p_val series."""
    return rng.choice(np.arange(0, target_dims * p_val, B), size=p_val, replace=False)


# the LV window gather (lotka_volterra_partial_batch_fix_theta.py:481-503) is statement for statement the FHN one
gather_feed_lv = gather_feed_fhn


# ----------------------------------------------------------------------------
# torch half: flow, ELBO, gradients, Adamax ("parity unpinned", see header)
# ----------------------------------------------------------------------------

def unpack_params(flat: torch.Tensor, layout) -> Dict[str, torch.Tensor]:
    return {k: flat[o:o + int(np.prod(s))].reshape(s) for k, (o, s) in layout.items()}


def _bn_affine(x, gamma, beta):
    """tf.layers.batch_normalization with training=False and never-updated moving stats
    (fitz_nag_NVP.py:93): gamma * (x - 0) / sqrt(1 + 1e-3) + beta."""
    return x * (gamma / math.sqrt(1.0 + 1e-3)) + beta


def flow_forward(cfg, P: Dict[str, torch.Tensor], eps: torch.Tensor, theta: torch.Tensor,
                 time_feats: torch.Tensor):
    """Flow_Stack.slp(): returns (x_final [p, L_F], logq [p]).

    eps [p, L0] is the base sample (AR.py:31-32, injected instead of sampled),
    theta [p, dtheta], time_feats [p, L0, Cf].
    """
    K, S, D = cfg.K, cfg.S, cfg.D
    x = eps
    # init_dist.slp: sum over the last S slots of log N(eps; 0, 1)            (AR.py:33-34)
    logq = (-0.5 * eps[:, -S:] ** 2 - 0.5 * LOG2PI).sum(dim=1)
    for i in range(cfg.F):
        if cfg.model in (3, 4, 5):
            # lotka_volterra_partial_batch_fix_theta.py:343-344,71-76 (lotka_volterra_partial.py:68-76,279-281): EVERY flow reads the whole window
            # (ts_feats = self.time_feats, no i*K slice); 3 x dense(50) + dense(feat_dims = L_i - 1), then the
            # [window position, unit] matrix is TRANSPOSED: unit m becomes the conv position, window position w a
            # conv input channel
            f = time_feats[:, :-1, :]
            for l in range(4):
                f = Fnn.elu(f @ P[f"f{i}.feat{l}.w"] + P[f"f{i}.feat{l}.b"])
            f = f.transpose(1, 2)
        else:
            ts = time_feats[:, i * K:, :]                                     # AR.py:192-193
            if cfg.feat_aug:                                                  # SV_dense.py:53
                f = torch.cat([ts[:, 1:, :], ts[:, 1:, :-2] - ts[:, :-1, :-2]], dim=2)
            else:
                f = ts[:, :-1, :]                                             # AR.py:53
            for l in range(4):                                                # AR.py:54-56
                f = Fnn.elu(f @ P[f"f{i}.feat{l}.w"] + P[f"f{i}.feat{l}.b"])
        inp = torch.cat([x[:, :-1, None], f], dim=2)                          # AR.py:58-59
        W = P[f"f{i}.conv.w"]                                                 # [K, Cin, Cout]
        A = Fnn.conv1d(inp.transpose(1, 2), W.permute(2, 1, 0), P[f"f{i}.conv.b"]).transpose(1, 2)  # AR.py:61-62
        b = theta
        for l in range(3):                                                    # AR.py:63-68
            b = b @ P[f"f{i}.th{l}.w"] + P[f"f{i}.th{l}.b"]
        h = Fnn.elu(A + b[:, None, :])                                        # AR.py:70-72
        for l in range(cfg.H):                                                # AR.py:74-76
            h = Fnn.elu(h @ P[f"f{i}.hid{l}.w"] + P[f"f{i}.hid{l}.b"])
            if cfg.bn:
                h = _bn_affine(h, P[f"f{i}.hid{l}.gamma"], P[f"f{i}.hid{l}.beta"])
        if D == 1:
            out = h @ P[f"f{i}.head.w"] + P[f"f{i}.head.b"]                   # AR.py:77-78
            mu, s = out[..., 0], out[..., 1]
            sigma = Fnn.softplus(s) + 1e-10                                   # AR.py:83
        else:
            # 1x1 conv with strides=2 -> heads at even conv positions          (fitz_nag_NVP.py:95-96)
            out = h[:, ::2, :] @ P[f"f{i}.head.w"] + P[f"f{i}.head.b"]
            mu_t, s_t = out[..., 0], out[..., 1]
            mu = torch.stack([torch.zeros_like(mu_t), mu_t], dim=2).reshape(x.shape[0], -1)       # :99-100
            sigma = torch.stack([torch.ones_like(s_t), Fnn.softplus(s_t) + 1e-10], dim=2).reshape(x.shape[0], -1)
        logq = logq - torch.log(sigma[:, -S:]).sum(dim=1)                     # AR.py:84,88
        x = x[:, K:] * sigma + mu                                             # AR.py:85
        if D == 2 and i < cfg.F - 1:
            # Permute: scatter_nd swaps each adjacent slot pair               (fitz_nag_NVP.py:145-153,205-211)
            x = x.reshape(x.shape[0], -1, 2).flip(2).reshape(x.shape[0], -1)
    return x, logq


def normal_logpdf(x, loc, scale):
    """tfd.Normal.log_prob."""
    scale = torch.as_tensor(scale, dtype=x.dtype)
    return -0.5 * ((x - loc) / scale) ** 2 - 0.5 * LOG2PI - torch.log(scale)


def elbo_terms(cfg, x_final: torch.Tensor, theta: torch.Tensor, time_feats: torch.Tensor,
               extra: Optional[Dict[str, torch.Tensor]] = None):
    """Per-row (sde_log_prob, obs_log_prob) and the latent path `lf_sample`."""
    B = cfg.B
    p = x_final.shape[0]
    if cfg.model == 0:      # AR.py:168-176
        lf = x_final                                                          # [p, B+1]
        obs_eval = time_feats[:, -B:, 0]                                      # AR.py:155
        w = time_feats[:, -B:, -1]
        obs_lp = (normal_logpdf(lf[:, 1:], obs_eval, cfg.obs_std) * w).sum(dim=1)
        head, tail = lf[:, :-1], lf[:, 1:]
        th = [theta[:, k:k + 1] for k in range(3)]
        sde_lp = normal_logpdf(tail, th[1] * head + th[0], torch.exp(th[2])).sum(dim=1)
        return sde_lp, obs_lp, lf
    if cfg.model == 1:      # fitz_nag_NVP.py:232-255
        lf = x_final.reshape(p, -1, 2).transpose(1, 2)                        # :282-283  [p,2,B+1]
        obs_eval = time_feats[:, -2 * B:, 0].reshape(p, -1, 2).transpose(1, 2)  # :216-217
        obs_lp = (normal_logpdf(lf[:, :, 1:], obs_eval, 0.1) * extra["bin_feed"]).reshape(p, -1).sum(dim=1)
        head, tail = lf[:, :, :-1], lf[:, :, 1:]
        diff = tail - head
        x1, x2 = head[:, 0, :], head[:, 1, :]
        th = [theta[:, k:k + 1] for k in range(5)]
        dt = cfg.dt
        d1 = torch.exp(th[0]) * (x1 - x1 ** 3 - x2 + th[1])
        d2 = th[2] * x1 - x2 + 1.4
        s1 = math.sqrt(dt) * torch.sqrt(torch.exp(th[3])).expand_as(x1)
        s2 = math.sqrt(dt) * torch.sqrt(torch.exp(th[4])).expand_as(x1)
        sde_lp = (normal_logpdf(diff[:, 0, :], dt * d1, s1) + normal_logpdf(diff[:, 1, :], dt * d2, s2)).sum(dim=1)
        return sde_lp, obs_lp, lf
    if cfg.model == 2:      # SV_dense.py:203-246
        lat = x_final * extra["mask"] + extra["shift"]
        lf = torch.stack([extra["dim_one"], lat], dim=1)                      # :245-246  [p,2,B+1]
        head, tail = lf[:, :, :-1], lf[:, :, 1:]
        diff = tail - head
        x1, x2 = head[:, 0, :], head[:, 1, :]
        th = [theta[:, k:k + 1] for k in range(4)]
        dt = cfg.dt
        d1 = th[0] * x1
        d2 = th[1] - torch.exp(th[2]) * x2
        s1 = math.sqrt(dt) * x1 * torch.exp(0.5 * x2)
        s2 = math.sqrt(dt) * torch.exp(th[3]).expand_as(x1)
        sde_lp = (normal_logpdf(diff[:, 0, :], dt * d1, s1) + normal_logpdf(diff[:, 1, :], dt * d2, s2)).sum(dim=1)
        return sde_lp, torch.zeros_like(sde_lp), lf
    if cfg.model == 3:      # lotka_volterra_partial_batch_fix_theta.py:265-332,346-371
        sde_lp, obs_lp, _, lf = lv_terms(cfg, x_final, theta, time_feats, extra)
        return sde_lp, obs_lp, lf
    if cfg.model == 4:      # lotka_volterra_partial.py:234-275,290-297
        sde_lp, obs_lp, _, lf = lvr_terms(cfg, x_final, theta, time_feats, extra)
        return sde_lp, obs_lp, lf
    raise ValueError("unknown model")


def lv_terms(cfg, x_final, theta, time_feats, extra, chain_on_transition=True):
    """Lotka-Volterra (fixed theta; with chain_on_transition=False the learned-theta batch script
    lotka_volterra_partial_batch.py:300-343, whose transition density is the plain bivariate normal of the state).  Returns (sde_log_prob incl. the x0 term, obs_log_prob, the log-det term that
    lf_log_prob receives on top of the flow's own logq, lf_sample [p,2,B+1]).

    Bijector conventions: Chain composes right to left; Softplus.ildj(y) = -log(1 - exp(-y));
    the chains' log-det is summed over the two components of each state ("event").  theta = softplus of the script's
    constants, constant across rows (ibid. :190)."""
    B, dt, p = cfg.B, cfg.dt, x_final.shape[0]
    neg = x_final.reshape(p, -1, 2).transpose(1, 2)                           # :350-351  [p,2,B+1]
    lf = (Fnn.softplus(neg) + 1.0) * extra["mask"] + extra["shift"]           # :355-358,367

    def ildj(y):            # of z -> 1 + softplus(z [- 1]) evaluated at y: -log(1 - exp(-(y - 1)))
        return -torch.log(-torch.expm1(-(y - 1.0)))

    def inv(y):             # inverse of Chain([Affine(+1), Softplus, Affine(-1)]): 1 + softplus^-1(y - 1)   (:307-314)
        return y + torch.log(-torch.expm1(-(y - 1.0)))      # = 1 + log(exp(y - 1) - 1), without the overflow
    extra_logq = ildj(lf[:, :, 1:]).reshape(p, -1).sum(dim=1)                 # :369-370
    th = [theta[:, k:k + 1] for k in range(4)]
    # observations (:266-272): y ~ 1 + softplus(N(x, theta3 x) - 1)
    obs_eval = time_feats[:, -2 * B:, 0].reshape(p, -1, 2).transpose(1, 2)
    loc = lf[:, :, 1:]
    y_lp = normal_logpdf(inv(obs_eval), loc, th[3][:, None, :] * loc) + ildj(obs_eval)
    obs_lp = (y_lp * extra["bin_feed"]).reshape(p, -1).sum(dim=1)
    # transitions (:274-314) between the states 1..B ("flow_head = lf_sample[:, :, 1:-1]")
    head, nxt = lf[:, :, 1:-1], lf[:, :, 2:]
    x1, x2 = head[:, 0, :], head[:, 1, :]
    a1 = th[0] * x1 - th[1] * x1 * x2
    a2 = th[1] * x1 * x2 - th[2] * x2
    ca = torch.sqrt(th[0] * x1 + th[1] * x1 * x2)
    cb = -th[1] * x1 * x2 / ca
    cc = torch.sqrt(th[1] * x1 * x2 + th[2] * x2 - cb ** 2)
    sq = math.sqrt(dt)
    L11, L21, L22 = sq * ca, sq * cb, sq * cc                                 # chol = sqrt(dt) [[a,0],[b,c]]
    tgt = inv(nxt) if chain_on_transition else nxt
    d1 = tgt[:, 0, :] - (x1 + dt * a1)
    d2 = tgt[:, 1, :] - (x2 + dt * a2)
    # Bivariate_Normal.normal_log_prob (:54-58): det = prod(diag(chol))^2, cov_inv = inverse(chol chol^T), both from the
    # un-jittered chol (the +1e-6 copy is stored but never used)
    w1 = d1 / L11
    w2 = (d2 - L21 * w1) / L22
    log_det = 2.0 * (torch.log(L11) + torch.log(L22))
    n_lp = -0.5 * log_det - 0.5 * (w1 ** 2 + w2 ** 2) - LOG2PI
    sde_lp = ((n_lp + ildj(nxt[:, 0, :]) + ildj(nxt[:, 1, :])) if chain_on_transition else n_lp).sum(dim=1)
    # p(x0) (:316-326): transformed diagonal Gaussian on the first retained state lf[:, :, 1]
    x0s = lf[:, :, 1]
    mean = torch.as_tensor(cfg.x0, dtype=x_final.dtype)
    x0_lp = (normal_logpdf(inv(x0s), mean, cfg.obs_std) + ildj(x0s)).sum(dim=1)   # x0_std rides in cfg.obs_std
    return sde_lp + x0_lp, obs_lp, extra_logq, lf


def lvb_theta_prior(theta, priors):
    """lotka_volterra_partial_batch.py:358-365: TransformedDistribution(MultivariateNormalDiag(mean, scale),
    Softplus(event_ndims=2)).log_prob(theta) = log N(softplus^-1(theta); mean, scale) - sum log(1 - exp(-theta)), the
    log-det summed over the components of a row."""
    mean = torch.as_tensor([m for m, _ in priors], dtype=theta.dtype)
    sd = torch.as_tensor([s for _, s in priors], dtype=theta.dtype)
    u = theta + torch.log(-torch.expm1(-theta))
    return (normal_logpdf(u, mean, sd) - torch.log(-torch.expm1(-theta))).sum(dim=1)


def pad_series_lvr(obs, time_till, x0, dt, T, target_dims, F, K, fw) -> Dict[str, object]:
    """lotka_volterra_partial.py:186-205 (flow_dims = 2)."""
    D = 2
    obs_flatten = np.reshape(obs, -1, 'F')
    store = []
    for i in range(0, fw * 5, 5):
        store.append(np.concatenate((np.zeros(F * K + D - i), obs_flatten, np.zeros(i)), axis=0))
    time_pad = np.concatenate((np.zeros(F * K + D), np.repeat(np.arange(dt, T + dt, dt), D)), axis=0)
    time_till_pad = np.reshape(np.repeat(np.arange(np.round((F * K + D) * (dt / D), 1), 0., -dt), D), (D, -1), 'F')
    return {
        "obs_pad_store": store,
        "time_pad": time_pad,
        "time_till": np.reshape(np.concatenate((time_till_pad, time_till), 1), -1, 'F'),
        "bin_feats": np.float32(np.concatenate((np.zeros(F * K + D), np.ones(target_dims * D)), axis=0)),
        "mask_vals": np.concatenate((np.zeros((2, 1)), np.ones((D, target_dims))), axis=1),
        "shift_vals": np.concatenate((np.expand_dims(x0, 1), np.zeros((D, target_dims))), axis=1),
    }


def lvr_terms(cfg, x_final, theta, time_feats, extra):
    """Lotka-Volterra, learned theta (lotka_volterra_partial.py).  Returns (sde_log_prob, obs_log_prob, the log-det
    term lf_log_prob receives on top of the flow's own logq, lf_sample [p,2,B+1]).

    :290-297  lf_sample = Softplus(event_ndims=2).forward(flow output [p,2,B+1]) * mask + shift;
              lf_log_prob += Softplus(event_ndims=2).inverse_log_det_jacobian(lf_sample[:, :, 1:])
              = sum over the last two axes of -log(1 - exp(-y))            (one value per row)
    :220-224  theta_eval = exp(theta sample), three rates
    :235      observations Normal(loc=obs_eval, scale=1).log_prob(lf_sample[:, :, 1:]) * bin_feed
    :237-262  Bivariate_Normal(mu = dt alpha(x_t), chol = sqrt(dt) sqrt_beta(x_t)).log_prob(x_{t+1} - x_t), all B steps
    :39-52    log_prob = -1/2 log det - 1/2 d^T Sigma^-1 d - log 2 pi, det = prod(diag(chol))^2"""
    B, dt, p = cfg.B, cfg.dt, x_final.shape[0]
    neg = x_final.reshape(p, -1, 2).transpose(1, 2)                           # :286-287  [p,2,B+1]
    lf = Fnn.softplus(neg) * extra["mask"] + extra["shift"]                   # :288-289
    extra_logq = (-torch.log(-torch.expm1(-lf[:, :, 1:]))).reshape(p, -1).sum(dim=1)      # :291-293
    th = [torch.exp(theta[:, k:k + 1]) for k in range(3)]
    obs_eval = time_feats[:, -2 * B:, 0].reshape(p, -1, 2).transpose(1, 2)
    obs_lp = (normal_logpdf(lf[:, :, 1:], obs_eval, 1.0) * extra["bin_feed"]).reshape(p, -1).sum(dim=1)
    head, tail = lf[:, :, :-1], lf[:, :, 1:]
    x1, x2 = head[:, 0, :], head[:, 1, :]
    a1 = th[0] * x1 - th[1] * x1 * x2
    a2 = th[1] * x1 * x2 - th[2] * x2
    ca = torch.sqrt(th[0] * x1 + th[1] * x1 * x2)
    cb = -th[1] * x1 * x2 / ca
    cc = torch.sqrt(th[1] * x1 * x2 + th[2] * x2 - cb ** 2)
    sq = math.sqrt(dt)
    L11, L21, L22 = sq * ca, sq * cb, sq * cc
    d1 = (tail[:, 0, :] - x1) - dt * a1
    d2 = (tail[:, 1, :] - x2) - dt * a2
    w1 = d1 / L11
    w2 = (d2 - L21 * w1) / L22
    log_det = 2.0 * (torch.log(L11) + torch.log(L22))
    sde_lp = (-0.5 * log_det - 0.5 * (w1 ** 2 + w2 ** 2) - LOG2PI).sum(dim=1)
    return sde_lp, obs_lp, extra_logq, lf


def objective(cfg, obj: int, P, eps, theta, time_feats, extra=None, path_target: float = 0.0):
    """Scalar the library differentiates, plus the per-row terms [p,4] = (sde, obs, logq, base_lp).

    obj 0: -sum_rows scale*(sde - logq + obs)   [ELBO minus the host-side prior - log q(theta); AR.py:184-185,228-229]
    obj 1: -sum_rows obs                         [AR.py:201-202]
    obj 2: sum (lf_sample - path_target)^2       [fitz_nag_NVP.py:288-289; SV_dense.py:251-252]
    """
    x_final, logq = flow_forward(cfg, P, eps, theta, time_feats)
    if cfg.model in (3, 4, 5):
        if cfg.model == 5:          # lotka_volterra_partial_batch.py: learned theta, no chain on the transition density
            sde, obs, extra_logq, lf = lv_terms(cfg, x_final, theta, time_feats, extra, chain_on_transition=False)
        else:
            sde, obs, extra_logq, lf = (lv_terms if cfg.model == 3 else lvr_terms)(cfg, x_final, theta, time_feats, extra)
        logq = logq + extra_logq                                              # lf_log_prob, LV fix-theta :369-370
    else:
        sde, obs, lf = elbo_terms(cfg, x_final, theta, time_feats, extra)
    base = (-0.5 * eps[:, -cfg.S:] ** 2 - 0.5 * LOG2PI).sum(dim=1)
    terms = torch.stack([sde, obs, logq, base], dim=1)
    if obj == 0:
        loss = -(cfg.scale * (sde - logq + obs)).sum()
    elif obj == 1:
        loss = -obs.sum()
    elif obj == 2:
        loss = ((lf - path_target) ** 2).sum()
    else:
        raise ValueError("objective")
    return loss, terms, lf, x_final


def step_reference(cfg, layout, flat_params: torch.Tensor, eps, theta, time_feats, obj=0, extra=None,
                   path_target: float = 0.0):
    """Forward + autograd gradients of `objective` w.r.t. the flat blob and theta."""
    fp = flat_params.detach().clone().requires_grad_(True)
    th = theta.detach().clone().requires_grad_(True)
    loss, terms, lf, x_final = objective(cfg, obj, unpack_params(fp, layout), eps, th, time_feats, extra, path_target)
    gp, gth = torch.autograd.grad(loss, [fp, th], allow_unused=True)
    if gp is None:
        gp = torch.zeros_like(fp)
    if gth is None:
        gth = torch.zeros_like(th)
    return {"loss": loss.detach(), "terms": terms.detach(), "lf": lf.detach(), "x_final": x_final.detach(),
            "grad_params": gp, "grad_theta": gth}


def adamax_step(w, g, m, v, lr, beta1, beta2=0.999, eps=1e-8, clip=None):
    """AR.py:230-234 (clip_by_global_norm over ALL gradients) + optimisers/adamax.py:51-57.

    `g` here is the gradient slice for these variables; pass `clip=(clip_norm, global_norm)` to
    apply TF's `g * clip_norm / max(global_norm, clip_norm)`.
    No bias correction; eps sits inside the max.  Returns (w, m, v) updated copies.
    """
    if clip is not None:
        clip_norm, gnorm = clip
        g = g * (clip_norm / max(gnorm, clip_norm))
    v = beta1 * v + (1.0 - beta1) * g
    m = torch.maximum(beta2 * m + eps, g.abs())
    w = w - lr * (v / m)
    return w, m, v


def glorot_init(layout, total, gen: torch.Generator, dtype=torch.float32) -> torch.Tensor:
    """TF defaults: Glorot-uniform kernels, zero biases, BN gamma=1 beta=0."""
    flat = torch.zeros(total, dtype=dtype)
    for name, (off, shape) in layout.items():
        n = int(np.prod(shape))
        if name.endswith(".w"):
            if len(shape) == 3:       # conv kernel [K, Cin, Cout]: fan_in = K*Cin, fan_out = K*Cout
                fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
            else:
                fan_in, fan_out = shape
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            flat[off:off + n] = (torch.rand(n, generator=gen, dtype=torch.float64) * 2 - 1).to(dtype) * lim
        elif name.endswith(".gamma"):
            flat[off:off + n] = 1.0
    return flat
