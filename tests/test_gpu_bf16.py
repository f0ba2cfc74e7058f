"""GPU parity of the bf16-split conv path (mode bit 2 of nma_set_tensor_cores): the three conv GEMMs - forward, data
gradient, weight gradient - as a 2-term bfloat16 split on tcgen05 kind::f16 (nma_tc.cuh, TcP<true>).

Bars: the bare contractions against fp64 within 4e-5 of the largest value (the split keeps 16 significand bits per
operand: ~1.5e-5 worst case per product, random in sign); the full step against the oracle within the same 1e-4 the
other paths are held to (BASELINE.json north_star), errors printed."""
import numpy as np
import pytest
import torch

from viforssms_b200.config import ar_config
from test_gpu_parity import _ar_case, _check_step, _engine, _rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("nacc,K,Q", [(2, 50, 1000), (1, 50, 300), (2, 7, 513), (2, 1, 256)])
def test_bf16_contraction_matches_fp64(mode, nacc, K, Q):
    from viforssms_b200.engine import tc_conv_raw
    g = torch.Generator().manual_seed(K * 7 + Q)
    x = torch.randn(Q, 56, generator=g) * torch.exp(torch.randn(Q, 1, generator=g))
    x[:, 51:] = 0.0
    w = torch.randn(K, 51, 50, generator=g) * 0.1
    got = tc_conv_raw(x.cuda(), w.cuda(), mode=mode + 2, nacc=nacc).cpu().double()
    xd, wd = x.double(), w.double()
    nout = Q - K + 1
    want = torch.zeros(nout, 64, dtype=torch.float64)
    for k in range(K):
        if mode == 0:
            want[:, :50] += xd[k:k + nout, :51] @ wd[k]
        else:
            want[:, :51] += xd[k:k + nout, :50] @ wd[K - 1 - k].T
    err = (got[:nout] - want).abs().max().item()
    scale = want.abs().max().item()
    print("bf16-split contraction: max abs err / max |value| = %.2e (K=%d, mode=%d)" % (err / scale, K, mode))
    assert err <= 4e-5 * scale, (err, scale)
    assert got[:nout, 51:].abs().max().item() == 0.0


@pytest.mark.parametrize("K,Q", [(50, 3000), (10, 200), (7, 129), (1, 64), (64, 40000)])
def test_bf16_weight_gradient_matches_fp64(K, Q):
    """MN-major operands straight from the conv operand layout, tap pairs as M = 128, periodic TMEM drains."""
    from viforssms_b200.engine import tc_wgrad_raw
    g = torch.Generator().manual_seed(K * 13 + Q)
    x = torch.randn(Q, 56, generator=g)
    da = torch.randn(Q, 56, generator=g) * torch.exp(torch.randn(Q, 1, generator=g))
    x[:, 51:] = 0.0
    da[:, 50:] = 0.0
    got = tc_wgrad_raw(x.cuda(), da.cuda(), K, bf16=True).cpu().double()
    nout = Q - K + 1
    xd, dd = x.double(), da.double()
    want = torch.stack([xd[k:k + nout, :51].T @ dd[:nout, :50] for k in range(K)])
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    print("bf16-split wgrad: max abs err / max |value| = %.2e (K=%d, Q=%d)" % (err / scale, K, Q))
    assert err <= 4e-5 * scale, (err, scale)


@pytest.mark.parametrize("shape", [
    dict(p=3, K=10, B=7, F=2, H=1, feat_window=3),
    dict(p=33, K=10, B=3, F=1, H=1, feat_window=1),
])
def test_bf16_step_parity_small(shape):
    cfg = ar_config(T=400, **shape)
    _check_step(cfg, 400, seed=3, tc=7)


def test_bf16_step_parity_ar_default():
    """configs[0] (p=50, K=50, B=50, 3 flows) with the conv GEMMs in the bf16 split."""
    worst = _check_step(ar_config(), 5000, seed=1, tc=7)
    assert worst < 5e-5


def test_bf16_mode_is_refused_where_the_kernels_do_not_cover_it():
    from viforssms_b200.engine import NMAEngine
    with pytest.raises(RuntimeError):
        NMAEngine(ar_config(p=4, K=20, B=13, F=2, H=3, feat_window=5, T=400), tensor_cores=7)   # three hidden layers
    with pytest.raises(RuntimeError):
        NMAEngine(ar_config(p=4, K=10, B=7, F=2, H=1, feat_window=3, T=400), tensor_cores=5)    # SIMT feature kernels


def test_switching_formats_on_one_handle_keeps_results():
    """3xTF32 -> bf16 split -> 3xTF32 on the same workspace: the operand buffers are cleared at each switch, so the
    third result equals the first bit for bit and the second agrees with them to the split's accuracy."""
    cfg = ar_config(p=16)
    arrays, idx, layout, params, eps, theta, _ = _ar_case(cfg, 5000, seed=4)
    dev = torch.device("cuda")
    eng = _engine(cfg, 3)
    eng.set_series(arrays)
    args = (params.to(dev), eps.to(dev), theta.to(dev), torch.from_numpy(idx).to(dev))
    a = {k: v.clone() for k, v in eng.elbo_fwd_bwd(*args).items()}
    eng.set_tensor_cores(7)
    assert eng.bf16_split
    b = {k: v.clone() for k, v in eng.elbo_fwd_bwd(*args).items()}
    eng.set_tensor_cores(3)
    c = eng.elbo_fwd_bwd(*args)
    torch.cuda.synchronize()
    assert torch.equal(a["terms"], c["terms"]) and torch.equal(a["lf"], c["lf"])
    assert _rel(c["grad_params"], a["grad_params"]) < 1e-6          # atomics: order-dependent in the last bits
    assert _rel(b["grad_params"], a["grad_params"]) < 5e-5
    assert _rel(b["terms"], a["terms"]) < 1e-5
