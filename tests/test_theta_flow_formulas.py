"""The device theta-posterior kernels (viforssms_b200/csrc/nma_theta_flow.cu: k_theta_flow_fwd / k_theta_flow_bwd),
transliterated to float64 numpy statement by statement, against the host autograd module (viforssms_b200/theta_flow.py):
parameter layout, masks, permutation direction, clip-with-gradient, every hand-derived gradient - under both gradient
semantics of the masked kernels (mask multiplied in the forward pass / TensorFlow's masked_dense, where masked entries
receive a gradient that the kernel constraint wipes after the update)."""
import numpy as np
import pytest
import torch

from viforssms_b200.theta_flow import ThetaFlow

H = 5


def _layout(d):
    sizes = [(d, H), (H, H), (H, H), (H, 2 * d)]
    return sizes, sum(a * b + b for a, b in sizes)


def _act(a, relu):
    return np.maximum(a, 0.0) if relu else np.where(a > 0, a, np.expm1(np.minimum(a, 0.0)))


def _dact(h, relu):
    return (h > 0).astype(np.float64) if relu else np.where(h > 0, 1.0, h + 1.0)


def _mlp(P, masks, d, relu, z):
    sizes, _ = _layout(d)
    hs, off = [z], 0
    for i, ((a, b), m) in enumerate(zip(sizes, masks)):
        W = P[off:off + a * b].reshape(a, b); off += a * b
        bias = P[off:off + b]; off += b
        pre = hs[-1] @ (W * m) + bias
        hs.append(pre if i == 3 else _act(pre, relu))
    return hs                      # [z, h1, h2, h3, out]


def kernel_fwd(params, masks, perms, z0, d, nb, relu, loc, scale):
    _, LP = _layout(d)
    z = z0.copy()
    lp = np.sum(-0.5 * ((z - loc) / scale) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(scale))
    for k in range(nb):
        out = _mlp(params[k * LP:(k + 1) * LP], masks, d, relu, z)[-1]
        zn = np.zeros(d)
        for j in range(d):
            ls = min(max(out[2 * j + 1], -5.0), 3.0)
            zn[j] = (z[j] - out[2 * j]) * np.exp(-ls)
            lp += ls
        z = np.array([zn[perms[k][j]] for j in range(d)]) if k < nb - 1 else zn
    return z, lp


def kernel_bwd(params, masks, perms, z0, d, nb, relu, g_theta, g_logq, mask_grad=False):
    sizes, LP = _layout(d)
    G = np.zeros_like(params)
    zin, z = [], z0.copy()
    for k in range(nb):
        zin.append(z.copy())
        out = _mlp(params[k * LP:(k + 1) * LP], masks, d, relu, z)[-1]
        zn = np.array([(z[j] - out[2 * j]) * np.exp(-min(max(out[2 * j + 1], -5.0), 3.0)) for j in range(d)])
        z = np.array([zn[perms[k][j]] for j in range(d)]) if k < nb - 1 else zn
    gz = g_theta.copy()
    for k in range(nb - 1, -1, -1):
        P = params[k * LP:(k + 1) * LP]
        if k < nb - 1:
            gzn = np.zeros(d)
            for j in range(d):
                gzn[perms[k][j]] += gz[j]
        else:
            gzn = gz.copy()
        z = zin[k]
        hs = _mlp(P, masks, d, relu, z)
        out = hs[-1]
        gout = np.zeros(2 * d)
        gz = np.zeros(d)
        for j in range(d):
            ls = min(max(out[2 * j + 1], -5.0), 3.0)
            e = np.exp(-ls)
            gz[j] = gzn[j] * e
            gout[2 * j] = -gzn[j] * e
            gout[2 * j + 1] = -gzn[j] * (z[j] - out[2 * j]) * e + g_logq
        # back through the masked MLP: offsets of (W, b) of the four sublayers inside the layer's block
        offs, off = [], 0
        for a, b in sizes:
            offs.append((off, off + a * b)); off += a * b + b
        g = gout
        for i in (3, 2, 1, 0):
            a, b = sizes[i]
            wo, bo = offs[i]
            W = P[wo:wo + a * b].reshape(a, b)
            go = g if i == 3 else g * _dact(hs[i + 1], relu)
            G[k * LP + bo:k * LP + bo + b] += go
            G[k * LP + wo:k * LP + wo + a * b] += (np.outer(hs[i], go) * (1.0 if mask_grad else masks[i])).reshape(-1)
            g = (W * masks[i]) @ go
        gz = gz + g
    return G, gz


@pytest.mark.parametrize("mask_grad", [False, True])
@pytest.mark.parametrize("d,nb,act", [(3, 5, "elu"), (5, 4, "elu"), (4, 4, "relu")])
def test_theta_flow_kernel_formulas_against_the_host_module(d, nb, act, mask_grad):
    np.random.seed(7)
    flow = ThetaFlow(d, nb, base_loc=1.5, base_scale=0.5, activation=act, tf_mask_grad=mask_grad)
    g = torch.Generator().manual_seed(3)
    flat = flow.init_values(g).double()
    flat = flat + 0.3 * torch.randn(flat.shape, generator=g, dtype=torch.float64) * (flat != 0)   # keep masked entries zero
    bias_mask = torch.zeros_like(flat)
    sizes, LP = _layout(d)
    assert flat.numel() == nb * LP == flow.n_params
    off = 0
    for _ in range(nb):
        for a, b in sizes:
            off += a * b
            bias_mask[off:off + b] = 1.0
            off += b
    flat = (flat + 0.2 * torch.randn(flat.shape, generator=g, dtype=torch.float64) * bias_mask).requires_grad_(True)
    flow.bind(flat)
    flow.masks = [m.double() for m in flow.masks]
    p = 6
    z0 = (1.5 + 0.5 * torch.randn(p, d, generator=g, dtype=torch.float64)).requires_grad_(True)
    theta, lp = flow.sample_and_log_prob(z0)
    g_theta = torch.randn(p, d, generator=g, dtype=torch.float64)
    g_logq = torch.randn(p, generator=g, dtype=torch.float64)
    loss = (theta * g_theta).sum() + (lp * g_logq).sum()
    gflat, gz0 = torch.autograd.grad(loss, [flat, z0])
    # what the kernels are handed: the flat parameters, the four masks, the permutations
    params = flat.detach().numpy()
    masks = [m for m in flow.masks_np]
    perms = [pm.tolist() for pm in flow.perms]
    G = np.zeros_like(params)
    for r in range(p):
        th, l = kernel_fwd(params, masks, perms, z0[r].detach().numpy(), d, nb, act == "relu", 1.5, 0.5)
        assert np.allclose(th, theta[r].detach().numpy(), rtol=1e-12, atol=1e-12)
        assert abs(l - lp[r].item()) <= 1e-12 * max(1.0, abs(lp[r].item()))
        Gr, gz = kernel_bwd(params, masks, perms, z0[r].detach().numpy(), d, nb, act == "relu",
                            g_theta[r].numpy(), g_logq[r].item(), mask_grad)
        G += Gr
        # z0 enters log q through the base density as well; the kernel returns only the flow part
        base = -(z0[r].detach().numpy() - 1.5) / 0.25 * g_logq[r].item()
        assert np.allclose(gz + base, gz0[r].numpy(), rtol=1e-9, atol=1e-10)
    assert np.allclose(G, gflat.numpy(), rtol=1e-9, atol=1e-10 * np.abs(gflat.numpy()).max())
    masked = (flow.mask_flat().numpy() == 0)
    assert masked.any() and (np.abs(G[masked]).max() > 0) == mask_grad      # the semantics differ exactly there
    # the kernel constraint: re-masking leaves the (already masked) variables untouched
    before = flat.detach().clone()
    flow.constrain()
    assert torch.equal(flat.detach(), before)
