"""Step parity in the regime bench.py measures (BASELINE.json configs[4]): a series of 10^8 steps generated on the
device, window starts near 0, 5*10^7 and 10^8 - 50 (the raw time channel reaches 1e8, AR.py:139-140, where float32 has
a spacing of 8), all three arithmetic modes, with the time-channel weights scaled as bench.py scales them and with the
plain Glorot values the reference's initialiser gives; and the bf16 split at bench-scale row counts (the weight gradient
reduces ~300 positions x 2048 rows per flow through truncating TMEM accumulation, periodic drains and fp32 atomics)."""
import numpy as np
import pytest
import torch

from oracle import nma_oracle as O
from viforssms_b200.config import ar_config, param_layout

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _per_variable(gp, ref, layout):
    gn_all = ref["grad_params"].norm().item()
    worst, worst_name = 0.0, ""
    for name, (off, shape) in layout.items():
        k = int(np.prod(shape))
        want = ref["grad_params"][off:off + k]
        err = (gp[off:off + k].double() - want).norm().item()
        rel = err / max(want.norm().item(), 1e-6 * gn_all)
        if rel > worst:
            worst, worst_name = rel, name
    return worst, worst_name


@pytest.fixture(scope="module")
def stepper_1e8():
    """The bench's own series: T = 10^8, generated on the device by the A12 / A13 kernels."""
    from viforssms_b200.trainer import ARStepper
    st = ARStepper(T=10 ** 8, rows=12, device=torch.device("cuda", 0), seed=1)
    yield st
    st.close()


IDX_1E8 = np.array([0, 50, 100, 150, 49_999_950, 50_000_000, 50_000_050, 76_543_200, 99_999_800, 99_999_850,
                    99_999_900, 99_999_950], dtype=np.int64)


def test_gather_at_1e8_is_the_reference_feed(stepper_1e8):
    """Window values against AR.py:135-150,267-288 evaluated directly: the time channel is the f32 cast of the integer
    step index, the pad indicator is 1 only left of the series, the look-ahead channels are the same series shifted."""
    st = stepper_1e8
    cfg = st.cfg
    fw, P = 10, cfg.F * cfg.K + 1
    tf, mask, shift = st.eng.gather(torch.from_numpy(IDX_1E8).to(st.device))
    tf = tf.cpu().numpy()
    for r, i0 in enumerate(IDX_1E8):
        q = i0 + np.arange(cfg.L0)                                  # padded positions of the window
        want_time = np.maximum(q - P, 0).astype(np.float64).astype(np.float32)
        assert np.array_equal(tf[r, :, fw + 1], want_time), r
        assert np.array_equal(tf[r, :, fw], (q < P).astype(np.float32)), r
        for c in range(1, fw):                                      # channel c at slot j = channel 0 at slot j + c
            assert np.array_equal(tf[r, :cfg.L0 - c, c], tf[r, c:, 0]), (r, c)
    assert tf[-1, :, fw + 1].max() == np.float32(10 ** 8 - 1 + cfg.L0 - P - 50 + 1 - 1) or tf[-1, :, fw + 1].max() >= 99_999_900


@pytest.mark.parametrize("tc", [0, 3, 7])
@pytest.mark.parametrize("unscaled", [False, True])
def test_step_parity_at_1e8(stepper_1e8, tc, unscaled):
    st = stepper_1e8
    cfg, dev = st.cfg, st.device
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(5)
    params = O.glorot_init(layout, n, g, torch.float32)
    for name, (off, shape) in layout.items():
        if name.endswith(".b"):
            k = int(np.prod(shape))
            params[off:off + k] = 0.05 * torch.randn(k, generator=g)
    if not unscaled:                    # bench.py / ARStepper: first-layer weights of the raw time channel times 10 / T
        for i in range(cfg.F):
            off, shape = layout[f"f{i}.feat0.w"]
            params[off:off + shape[0] * shape[1]].reshape(shape)[11, :] *= 10.0 / 10 ** 8
    eps = torch.randn(cfg.p, cfg.L0, generator=g)
    theta = torch.stack([torch.randn(cfg.p, generator=g) * 0.5 + 4.0, torch.randn(cfg.p, generator=g) * 0.1 + 0.5,
                         torch.randn(cfg.p, generator=g) * 0.2 + 1.0], dim=1).float()
    idx = torch.from_numpy(IDX_1E8).to(dev)
    st.eng.set_tensor_cores(tc)
    tf, _, _ = st.eng.gather(idx)
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf.cpu().double())
    out = st.eng.elbo_fwd_bwd(params.to(dev), eps.to(dev), theta.to(dev), idx)
    torch.cuda.synchronize()
    terms = out["terms"].cpu().double()
    terr = max(((terms[:, k] - ref["terms"][:, k]).abs().max() / max(1.0, ref["terms"][:, k].abs().max().item())).item()
               for k in range(4))
    lferr = ((out["lf"].cpu().double() - ref["x_final"]).norm() / ref["x_final"].norm()).item()
    worst, wname = _per_variable(out["grad_params"].cpu(), ref, layout)
    gerr = ((out["grad_params"].cpu().double() - ref["grad_params"]).norm() / ref["grad_params"].norm()).item()
    print("T=1e8 step parity (tc=%d, %s time weights): terms %.2e, path %.2e, all-grad %.2e, worst variable %.2e (%s)"
          % (tc, "Glorot" if unscaled else "scaled", terr, lferr, gerr, worst, wname))
    assert torch.isfinite(terms).all()
    assert terr <= RTOL and lferr <= RTOL and gerr <= RTOL and worst <= RTOL


@pytest.mark.parametrize("rows", [2048, 4096])
def test_bf16_split_parity_at_bench_scale_rows(rows):
    """The headline mode (conv GEMMs in the 2-term bf16 split) against the fp64 oracle at thousands of rows, per variable."""
    from test_gpu_parity import _ar_case, _engine
    cfg = ar_config(p=rows)
    arrays, idx, layout, params, eps, theta, tf32 = _ar_case(cfg, 5000, seed=rows)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf32.double())
    dev = torch.device("cuda")
    res = {}
    for tc in (7, 3):
        eng = _engine(cfg, tc)
        eng.set_series(arrays)
        out = eng.elbo_fwd_bwd(params.to(dev), eps.to(dev), theta.to(dev), torch.from_numpy(idx).to(dev))
        torch.cuda.synchronize()
        terms = out["terms"].cpu().double()
        terr = max(((terms[:, k] - ref["terms"][:, k]).abs().max() / max(1.0, ref["terms"][:, k].abs().max().item())).item()
                   for k in range(4))
        worst, wname = _per_variable(out["grad_params"].cpu(), ref, layout)
        gerr = ((out["grad_params"].cpu().double() - ref["grad_params"]).norm() / ref["grad_params"].norm()).item()
        res[tc] = (terr, gerr, worst, wname)
        eng.close()
    for tc, (terr, gerr, worst, wname) in res.items():
        print("%d rows, tc=%d: terms %.2e, all-grad %.2e, worst variable %.2e (%s)" % (rows, tc, terr, gerr, worst, wname))
    for tc, (terr, gerr, worst, wname) in res.items():
        assert terr <= RTOL and gerr <= RTOL and worst <= RTOL, (tc, terr, gerr, worst, wname)
