"""The hand-derived gradients of the learned-theta Lotka-Volterra ELBO branch (k_elbo, NMA_MODEL_LVR in
viforssms_b200/csrc/nma_elbo.cu), transliterated to float64 numpy statement by statement, against autograd of the
oracle's lvr_terms (which is pinned to lotka_volterra_partial.py's own classes, tests/test_step_golden_models.py).
The kernel branch was written after the round's GPU budget was spent; this is its CPU pre-flight check."""
import math

import numpy as np
import torch

from oracle import nma_oracle as O
from viforssms_b200.config import lvr_config

LOG2PI = math.log(2 * math.pi)


def softplus(v):
    return np.maximum(v, 0) + np.log1p(np.exp(-abs(v)))


def sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def kernel_row(x, th, y, w, i0, B, dt, scale, x0, objective, path_target):
    """One row of the LVR branch: returns (sde, obs, lq_extra, dx [2(B+1)], gth [3], lf [2(B+1)])."""
    t0, t1, t2 = np.exp(th)
    c_sde = c_obs = c_sq = 0.0
    if objective == 0:
        c_sde = c_obs = -scale
    elif objective == 1:
        c_obs = -1.0
    else:
        c_sq = 1.0
    cq = scale if objective == 0 else 0.0
    sde = obs = lq = 0.0
    gth = np.zeros(3)
    dx = np.zeros(2 * (B + 1))
    lf = np.zeros(2 * (B + 1))
    for j in range(B + 1):
        first = (i0 + j) == 0
        z1, z2 = x[2 * j], x[2 * j + 1]
        u1 = x0[0] if first else softplus(z1)
        u2 = x0[1] if first else softplus(z2)
        g1 = g2 = h1 = h2 = 0.0
        if j >= 1:
            lq += softplus(-z1) + softplus(-z2)
            h1 += cq * (-sigmoid(-z1))
            h2 += cq * (-sigmoid(-z2))
            y1, y2, w1, w2 = y[0, j - 1], y[1, j - 1], w[0, j - 1], w[1, j - 1]
            obs += w1 * (-0.5 * (u1 - y1) ** 2 - 0.5 * LOG2PI) + w2 * (-0.5 * (u2 - y2) ** 2 - 0.5 * LOG2PI)
            g1 += c_obs * (-(u1 - y1)) * w1
            g2 += c_obs * (-(u2 - y2)) * w2
            pf = (i0 + j - 1) == 0
            p1 = x0[0] if pf else softplus(x[2 * j - 2])
            p2 = x0[1] if pf else softplus(x[2 * j - 1])
            s11, s12, s22 = t0 * p1 + t1 * p1 * p2, -t1 * p1 * p2, t1 * p1 * p2 + t2 * p2
            Dd = s11 * s22 - s12 * s12
            d1 = (u1 - p1) - dt * (t0 * p1 - t1 * p1 * p2)
            d2 = (u2 - p2) - dt * (t1 * p1 * p2 - t2 * p2)
            g1 += c_sde * (-(s22 * d1 - s12 * d2) / (dt * Dd))
            g2 += c_sde * (-(-s12 * d1 + s11 * d2) / (dt * Dd))
        if j < B:
            n1, n2 = softplus(x[2 * j + 2]), softplus(x[2 * j + 3])
            s11, s12, s22 = t0 * u1 + t1 * u1 * u2, -t1 * u1 * u2, t1 * u1 * u2 + t2 * u2
            Dd = s11 * s22 - s12 * s12
            d1 = (n1 - u1) - dt * (t0 * u1 - t1 * u1 * u2)
            d2 = (n2 - u2) - dt * (t1 * u1 * u2 - t2 * u2)
            Qf = s22 * d1 * d1 - 2 * s12 * d1 * d2 + s11 * d2 * d2
            sde += -math.log(dt) - 0.5 * math.log(Dd) - 0.5 * Qf / (dt * Dd) - LOG2PI
            gm1, gm2 = (s22 * d1 - s12 * d2) / (dt * Dd), (-s12 * d1 + s11 * d2) / (dt * Dd)
            e1 = gm1 * (1 + dt * (t0 - t1 * u2)) + gm2 * (dt * t1 * u2)
            e2 = gm1 * (-dt * t1 * u1) + gm2 * (1 + dt * (t1 * u1 - t2))
            i2 = 1.0 / (2 * dt * Dd * Dd)
            L11 = -0.5 * s22 / Dd - (d2 * d2 * Dd - Qf * s22) * i2
            L22 = -0.5 * s11 / Dd - (d1 * d1 * Dd - Qf * s11) * i2
            L12 = s12 / Dd - (-2 * d1 * d2 * Dd + 2 * Qf * s12) * i2
            e1 += L11 * (t0 + t1 * u2) + L12 * (-t1 * u2) + L22 * (t1 * u2)
            e2 += L11 * (t1 * u1) + L12 * (-t1 * u1) + L22 * (t1 * u1 + t2)
            g1 += c_sde * e1
            g2 += c_sde * e2
            uu = u1 * u2
            gth[0] += t0 * (gm1 * dt * u1 + L11 * u1)
            gth[1] += t1 * (dt * uu * (gm2 - gm1) + (L11 - L12 + L22) * uu)
            gth[2] += t2 * (-gm2 * dt * u2 + L22 * u2)
        g1 += c_sq * 2 * (u1 - path_target)
        g2 += c_sq * 2 * (u2 - path_target)
        dx[2 * j] = 0.0 if first else g1 * sigmoid(z1) + h1
        dx[2 * j + 1] = 0.0 if first else g2 * sigmoid(z2) + h2
        lf[2 * j], lf[2 * j + 1] = u1, u2
    return sde, obs, lq, dx, c_sde * gth, lf


def test_lvr_kernel_formulas_against_autograd():
    rs = np.random.RandomState(3)
    B, dt, p = 9, 0.1, 4
    cfg = lvr_config(p=p, K=4, B=B, F=2, H=1, feat_window=2, target_dims=90, dt=dt, x0=(100.0, 90.0))
    i0s = np.array([0, 9, 27, 45])
    x = rs.normal(4.0, 1.0, size=(p, 2 * (B + 1)))                     # raw flow output: states softplus(z) ~ 2..6
    th = np.log(np.array([0.45, 0.05, 0.3]))[None, :] + 0.1 * rs.standard_normal((p, 3))
    y = rs.normal(4.0, 2.0, size=(p, 2, B))
    w = (rs.uniform(size=(p, 2, B)) < 0.4).astype(np.float64)
    mask = np.ones((p, 2, B + 1)); shift = np.zeros((p, 2, B + 1))
    mask[0, :, 0] = 0.0
    shift[0, :, 0] = cfg.x0
    # time_feats with the observations in channel 0 of the last 2B slots, interleaved as the script lays them out
    tfeats = np.zeros((p, cfg.L0, cfg.Cf))
    tfeats[:, -2 * B:, 0] = y.transpose(0, 2, 1).reshape(p, -1)
    for objective, target in ((0, 0.0), (2, 75.0)):
        xt = torch.tensor(x, requires_grad=True)
        tht = torch.tensor(th, requires_grad=True)
        extra = {"mask": torch.tensor(mask), "shift": torch.tensor(shift), "bin_feed": torch.tensor(w)}
        sde, obs, lq, lf = O.lvr_terms(cfg, xt, tht, torch.tensor(tfeats), extra)
        if objective == 0:
            loss = -(cfg.scale * (sde - lq + obs)).sum()              # logq = (flow part, constant here) + lq
        else:
            loss = ((lf - target) ** 2).sum()
        gx, gth = torch.autograd.grad(loss, [xt, tht], allow_unused=True)
        for r in range(p):
            ks, ko, kl, kdx, kgth, klf = kernel_row(x[r], th[r], y[r], w[r], int(i0s[r]), B, dt, cfg.scale, cfg.x0,
                                                    objective, target)
            assert abs(ks - sde[r].item()) <= 1e-9 * max(1.0, abs(sde[r].item()))
            assert abs(ko - obs[r].item()) <= 1e-9 * max(1.0, abs(obs[r].item()))
            assert abs(kl - lq[r].item()) <= 1e-9 * max(1.0, abs(lq[r].item()))
            assert np.allclose(klf.reshape(-1, 2).T, lf[r].detach().numpy(), rtol=1e-12, atol=1e-12)
            assert np.allclose(kdx, gx[r].numpy(), rtol=1e-8, atol=1e-8 * np.abs(gx[r].numpy()).max())
            if objective == 0:
                assert np.allclose(kgth, gth[r].numpy(), rtol=1e-8, atol=1e-8 * np.abs(gth[r].numpy()).max())
