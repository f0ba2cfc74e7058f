"""GPU leg of tests/test_step_golden.py: the CUDA path, through the C-ABI, against values produced by the
reference's own classes (tests/golden/ar_step_golden.npz; see make_golden_step.py for how they were made).
Bar: 1e-4 relative (BASELINE.json north_star), per ELBO term and per gradient variable."""
import numpy as np
import pytest
import torch

from test_step_golden import G, case_inputs  # noqa: F401  (G is a fixture)

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _run(cfg, params, eps, theta, idx, arrays, tc):
    from viforssms_b200.engine import NMAEngine
    eng = NMAEngine(cfg, tensor_cores=tc)
    eng.set_series(arrays)
    dev = torch.device("cuda")
    out = eng.elbo_fwd_bwd(params.to(dev), eps.to(dev), theta.to(dev), torch.from_numpy(idx).to(dev))
    torch.cuda.synchronize()
    return eng, out


# 0: FP32 SIMT kernels, 3: tcgen05 3xTF32 (library default), 7: conv GEMMs in the bf16 split
@pytest.mark.parametrize("tc", [0, 3, 7])
@pytest.mark.parametrize("case", ["small", "full"])
def test_cuda_step_matches_the_reference_classes(G, case, tc):
    cfg, layout, n, params, eps, theta, idx, tf32, arrays = case_inputs(G, case)
    eng, out = _run(cfg, params, eps, theta, idx, arrays, tc)
    t = out["terms"].cpu().double().numpy()
    for k, name in enumerate(("sde", "obs", "logq")):
        want = G["%s_%s" % (case, name)]
        assert np.abs(t[:, k] - want).max() <= RTOL * max(1.0, np.abs(want).max()), name
    lf = out["lf"].cpu().double().numpy()
    want = G[case + "_lf_sample"]
    assert np.linalg.norm(lf - want) <= RTOL * np.linalg.norm(want)
    gth = out["grad_theta"].cpu().double().numpy()
    want = G[case + "_grad_theta"]
    assert np.linalg.norm(gth - want) <= RTOL * np.linalg.norm(want)
    gp = out["grad_params"].cpu().double().numpy()
    worst = 0.0
    if case == "small":
        want = G["small_grad"]
        gn = np.linalg.norm(want)
        for name, (off, shape) in layout.items():
            k = int(np.prod(shape))
            err = np.linalg.norm(gp[off:off + k] - want[off:off + k])
            scale = max(np.linalg.norm(want[off:off + k]), 1e-6 * gn)
            worst = max(worst, err / scale)
            assert err <= RTOL * scale, name
    else:
        heads = G["full_grad_heads"]
        pos = 0
        gn = float(G["full_global_norm"])
        for nm, want_norm in zip([str(s) for s in G["full_var_names"]], G["full_grad_norms"]):
            off, shape = layout[nm]
            k = int(np.prod(shape))
            scale = max(want_norm, 1e-6 * gn)
            assert abs(np.linalg.norm(gp[off:off + k]) - want_norm) <= RTOL * scale, nm
            h = min(k, 16)
            err = np.linalg.norm(gp[off:off + h] - heads[pos:pos + h])
            worst = max(worst, err / max(np.linalg.norm(heads[pos:pos + h]), 1e-6 * gn))
            assert err <= RTOL * max(np.linalg.norm(heads[pos:pos + h]), 1e-6 * gn), nm
            pos += h
    gnorm = float(G[case + "_global_norm"])
    assert abs(np.linalg.norm(gp) - gnorm) <= RTOL * gnorm
    print("CUDA vs reference classes (%s, mode %d): worst per-variable gradient error %.2e" % (case, tc, worst))


def test_cuda_adamax_matches_the_reference_optimizer(G):
    """nma_adamax_step (global-norm clip + Adamax, AR.py:226-234 / optimisers/adamax.py:42-58) on the golden gradient,
    from zero slots, against the variables the reference's AdamaxOptimizer left behind."""
    cfg, layout, n, params, eps, theta, idx, tf32, arrays = case_inputs(G, "small")
    from viforssms_b200.engine import NMAEngine
    eng = NMAEngine(cfg)
    dev = torch.device("cuda")
    lr, clip = (float(v) for v in G["small_lr_clip"])
    w = params.to(dev).clone()
    g = torch.from_numpy(G["small_grad"]).float().to(dev)
    m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    norm = eng.adamax_step(w, g, m, v, lr, 0.95, clip=clip)
    torch.cuda.synchronize()
    gn = float(G["small_global_norm"])
    assert abs(norm.item() - gn) <= 1e-5 * gn
    want = torch.from_numpy(G["small_params_after"])
    assert torch.allclose(w.cpu(), want, rtol=0, atol=3e-7)
    # the step is sign(g) * lr * (1 - beta1) wherever |g| > 1e-8 (first step from zero slots)
    big = torch.from_numpy(np.abs(G["small_grad"]) > 1e-3)
    upd = (w.cpu().double() - params.double())[big]
    assert (upd.abs() - lr * 0.05).abs().max().item() <= 1e-4 * lr


# ---------------------------------------------------------------------------------------------------
# FitzHugh-Nagumo and stochastic volatility: the CUDA path against the reference's own classes
# (tests/golden/models_step_golden.npz, tests/golden/make_golden_step_models.py)
# ---------------------------------------------------------------------------------------------------
from test_step_golden_models import GM, check_grads, fhn_inputs, sv_inputs  # noqa: E402,F401


def _check_model(GM, prefix, inputs, term_keys, tc, objective=0, path_target=0.0, gprefix=None):
    from viforssms_b200.engine import NMAEngine
    cfg, layout, n, params, eps, theta, idx, tf, extra, arrays = inputs
    eng = NMAEngine(cfg, tensor_cores=tc)
    eng.set_series(arrays)
    dev = torch.device("cuda")
    out = eng.elbo_fwd_bwd(params.to(dev), eps.to(dev), theta.to(dev), torch.from_numpy(np.asarray(idx)).to(dev),
                           objective=objective, path_target=path_target)
    torch.cuda.synchronize()
    gp = out["grad_params"].cpu().double().numpy()
    worst = check_grads(gp, layout, GM, gprefix or prefix, RTOL, 1e-6)
    if objective == 0:
        t = out["terms"].cpu().double().numpy()
        for k, key in term_keys:
            want = GM[prefix + key]
            assert np.abs(t[:, k] - want).max() <= RTOL * max(1.0, np.abs(want).max()), key
        gth = out["grad_theta"].cpu().double().numpy()
        want = GM[prefix + "grad_theta"]
        assert np.linalg.norm(gth - want) <= RTOL * np.linalg.norm(want)
    return out, worst


def test_cuda_fhn_step_matches_the_reference_classes(GM):
    inputs = fhn_inputs(GM)
    out, worst = _check_model(GM, "fhn_", inputs, [(0, "sde"), (1, "obs_lp"), (2, "logq")], None)
    p, B = inputs[0].p, inputs[0].B
    lf = out["lf"].cpu().double().numpy().reshape(p, -1, 2).transpose(0, 2, 1)      # fitz_nag_NVP.py:282-283
    want = GM["fhn_lf_sample"]
    assert np.linalg.norm(lf - want) <= RTOL * np.linalg.norm(want)
    _, worst2 = _check_model(GM, "fhn_", inputs, [], None, objective=2, path_target=0.0, gprefix="fhn_pre_")
    print("CUDA vs reference classes (FHN): worst gradient slice error %.2e (ELBO), %.2e (pre-training)" % (worst, worst2))


@pytest.mark.parametrize("tc", [0, 3])
def test_cuda_sv_step_matches_the_reference_classes(GM, tc):
    inputs = sv_inputs(GM)
    out, worst = _check_model(GM, "sv_", inputs, [(0, "sde"), (2, "logq")], tc)
    _, worst2 = _check_model(GM, "sv_", inputs, [], tc, objective=2, path_target=-7.0, gprefix="sv_pre_")
    print("CUDA vs reference classes (SV, mode %d): worst gradient slice error %.2e (ELBO), %.2e (pre-training)" % (tc, worst, worst2))
