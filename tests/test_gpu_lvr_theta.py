"""GPU parity tests of the learned-theta Lotka-Volterra model (lotka_volterra_partial.py, NMA_MODEL_LVR), the fixed-theta
one against the reference-classes fixture, and the theta posterior on the device (nma_theta_flow_fwd / _bwd) against the
host autograd module.  (First run on a B200 at the start of round 2: all green, profiles/r02_first_gpu_run.log.)
"""
import os

import numpy as np
import pytest
import torch

from oracle import nma_oracle as O
from test_step_golden_models import GM, check_grads, lvr_inputs  # noqa: F401

pytestmark = [pytest.mark.gpu]
RTOL = 1e-4


@pytest.mark.parametrize("objective,target,gprefix", [(0, 0.0, "lvr_"), (2, 75.0, "lvr_pre_")])
def test_cuda_lvr_step_matches_the_reference_classes(GM, objective, target, gprefix):
    from viforssms_b200.engine import NMAEngine
    cfg, layout, n, params, eps, theta, idx, tf, extra, arrays = lvr_inputs(GM)
    eng = NMAEngine(cfg)
    eng.set_series(arrays)
    got_tf, got_mask, got_shift = eng.gather(idx)
    assert np.array_equal(got_tf.cpu().numpy(), tf.numpy().astype(np.float32))
    assert np.array_equal(got_mask.cpu().numpy(), extra["mask"].numpy().astype(np.float32))
    assert np.array_equal(got_shift.cpu().numpy(), extra["shift"].numpy().astype(np.float32))
    dev = torch.device("cuda")
    out = eng.elbo_fwd_bwd(params.to(dev), eps.to(dev), theta.to(dev), torch.from_numpy(np.asarray(idx)).to(dev),
                           objective=objective, path_target=target)
    torch.cuda.synchronize()
    worst = check_grads(out["grad_params"].cpu().double().numpy(), layout, GM, gprefix, RTOL, 1e-6)
    ref = O.step_reference(cfg, layout, params.double(), eps.double(), theta.double(), tf, obj=objective, extra=extra,
                           path_target=target)
    gth = out["grad_theta"].cpu().double().numpy()
    want = ref["grad_theta"].numpy()
    assert np.linalg.norm(gth - want) <= RTOL * max(np.linalg.norm(want), 1e-6 * float(GM[gprefix + "global_norm"]))
    if objective == 0:
        t = out["terms"].cpu().double().numpy()
        for k, key in ((0, "sde"), (1, "obs_lp"), (2, "logq")):
            w = GM["lvr_" + key]
            assert np.abs(t[:, k] - w).max() <= RTOL * max(1.0, np.abs(w).max()), key
        lf = out["lf"].cpu().double().numpy().reshape(cfg.p, -1, 2).transpose(0, 2, 1)
        assert np.linalg.norm(lf - GM["lvr_lf_sample"]) <= RTOL * np.linalg.norm(GM["lvr_lf_sample"])
        assert np.linalg.norm(gth - GM["lvr_grad_theta"]) <= RTOL * np.linalg.norm(GM["lvr_grad_theta"])
    print("CUDA vs reference classes (LV learned theta, objective %d): worst gradient slice error %.2e" % (objective, worst))


def test_cuda_lv_fixed_theta_matches_the_reference_classes(GM):
    """The (hardware-verified) fixed-theta LV path against the reference-classes fixture under the per-state reading;
    parked here only because the test itself was written without a GPU at hand."""
    from test_step_golden_models import lv_fixed_inputs
    from viforssms_b200.engine import NMAEngine
    cfg, layout, n, params, eps, theta, idx, tf, extra, arrays = lv_fixed_inputs(GM)
    eng = NMAEngine(cfg)
    eng.set_series(arrays)
    dev = torch.device("cuda")
    out = eng.elbo_fwd_bwd(params.to(dev), eps.to(dev), theta.to(dev), torch.from_numpy(np.asarray(idx)).to(dev))
    torch.cuda.synchronize()
    pre = "lvf_per_state_"
    worst = check_grads(out["grad_params"].cpu().double().numpy(), layout, GM, pre, RTOL, 1e-6)
    t = out["terms"].cpu().double().numpy()
    for k, key in ((0, "sde"), (1, "obs_lp"), (2, "logq")):
        w = GM[pre + key]
        assert np.abs(t[:, k] - w).max() <= RTOL * max(1.0, np.abs(w).max()), key
    lf = out["lf"].cpu().double().numpy().reshape(cfg.p, -1, 2).transpose(0, 2, 1)
    assert np.linalg.norm(lf - GM[pre + "lf_sample"]) <= RTOL * np.linalg.norm(GM[pre + "lf_sample"])
    print("CUDA vs reference classes (LV fixed theta, per-state reading): worst gradient slice error %.2e" % worst)


@pytest.mark.parametrize("mask_grad", [False, True])
@pytest.mark.parametrize("d,nb,act", [(3, 5, "elu"), (5, 4, "elu"), (4, 4, "relu")])
def test_device_theta_flow_matches_the_host_module(d, nb, act, mask_grad):
    """nma_theta_flow_fwd / _bwd (two launches) against the host autograd module they are to replace."""
    from viforssms_b200.engine import DeviceThetaFlow
    from viforssms_b200.theta_flow import ThetaFlow
    np.random.seed(7)
    dev = torch.device("cuda")
    flow = ThetaFlow(d, nb, base_loc=1.5, base_scale=0.5, activation=act, tf_mask_grad=mask_grad)
    g = torch.Generator().manual_seed(3)
    flat = flow.init_values(g)
    flat = (flat + 0.3 * torch.randn(flat.shape, generator=g) * (flat != 0)).to(dev).requires_grad_(True)
    flow.bind(flat)
    p = 1000
    z0 = (1.5 + 0.5 * torch.randn(p, d, generator=g)).to(dev)
    theta, lp = flow.sample_and_log_prob(z0)
    g_theta = torch.randn(p, d, generator=g).to(dev)
    g_logq = torch.randn(p, generator=g).to(dev)
    ((theta * g_theta).sum() + (lp * g_logq).sum()).backward()
    dflow = DeviceThetaFlow(flow, dev)
    th2, lp2 = dflow.forward(flat.detach(), z0)
    gp = torch.zeros_like(flat)
    dflow.backward(flat.detach(), z0, g_theta, g_logq, gp)
    torch.cuda.synchronize()
    assert torch.allclose(th2, theta.detach(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(lp2, lp.detach(), rtol=1e-5, atol=1e-4)
    err = (gp - flat.grad).norm().item() / flat.grad.norm().item()
    print("device theta flow (d=%d, nb=%d, %s): gradient rel err %.2e" % (d, nb, act, err))
    assert err < 1e-4


@pytest.mark.parametrize("mask_grad", [False, True])
def test_ar_stepper_train_step_matches_the_host_theta_path(mask_grad):
    """One training iteration as ONE nma_train_step call (in-library noise, device theta posterior, clip + Adamax)
    against the same iteration composed on the host (autograd theta posterior, nma_elbo_fwd_bwd, nma_adamax_step) with
    the noise the library drew injected: gradient blobs and updated variables must agree, under both gradient semantics
    of the masked kernels."""
    from viforssms_b200.trainer import ARStepper
    dev = torch.device("cuda", 0)
    a = ARStepper(T=20000, rows=64, device=dev, seed=5, device_theta=True, tf_mask_grad=mask_grad)
    a._step(a.idx_dev)
    torch.cuda.synchronize()
    buf = a.eng.step_buffers(a.rows)
    assert a.eng.draw_counter() == 1
    b = ARStepper(T=20000, rows=64, device=dev, seed=5, device_theta=False, tf_mask_grad=mask_grad)
    assert torch.equal(a.idx_dev, b.idx_dev)
    elbo_b = b._step_host_theta(b.idx_dev, z0=buf["z0"], eps=buf["eps"])
    torch.cuda.synchronize()
    g0, g1 = b.grad, a.grad
    err = (g1 - g0).norm().item() / g0.norm().item()
    tail = (g1[-580:] - g0[-580:]).norm().item() / g0[-580:].norm().item()
    print("nma_train_step vs host composition (mask_grad=%s): gradient rel err %.2e (flow variables alone %.2e)"
          % (mask_grad, err, tail))
    assert err < 1e-5 and tail < 1e-4
    masked = (a.flow.mask_flat() == 0).to(dev)
    assert (g1[-580:][masked].abs().max().item() > 0) == mask_grad
    assert torch.allclose(a.blob, b.blob, rtol=0, atol=1e-5)
    assert torch.all(a.blob[-580:][masked] == 0)                   # the kernel constraint
    assert abs(a.scalars[0].item() - elbo_b.item()) <= 1e-4 * abs(elbo_b.item())
    a.close(); b.close()


def test_cuda_lv_batch_step_matches_the_reference_classes(GM):
    """lotka_volterra_partial_batch.py (learned softplus-theta, plain bivariate transition density, the first p_val states
    pinned): gather, terms, path, every gradient and d/dtheta against the script's own classes (tests/golden)."""
    from test_step_golden_models import lvb_inputs
    from viforssms_b200.engine import NMAEngine
    cfg, layout, n, params, eps, theta, idx, tf, extra, arrays = lvb_inputs(GM)
    eng = NMAEngine(cfg)
    eng.set_series(arrays)
    got_tf, got_mask, got_shift = eng.gather(idx)
    assert np.array_equal(got_tf.cpu().numpy(), tf.numpy().astype(np.float32))
    assert np.array_equal(got_mask.cpu().numpy(), extra["mask"].numpy().astype(np.float32))
    assert np.array_equal(got_shift.cpu().numpy(), extra["shift"].numpy().astype(np.float32))
    dev = torch.device("cuda")
    out = eng.elbo_fwd_bwd(params.to(dev), eps.float().to(dev), theta.float().to(dev), torch.from_numpy(np.asarray(idx)).to(dev))
    torch.cuda.synchronize()
    worst = check_grads(out["grad_params"].cpu().double().numpy(), layout, GM, "lvb_", RTOL, 1e-6)
    t = out["terms"].cpu().double().numpy()
    for k, key in ((0, "sde"), (1, "obs_lp"), (2, "logq")):
        w = GM["lvb_" + key]
        assert np.abs(t[:, k] - w).max() <= RTOL * max(1.0, np.abs(w).max()), key
    lf = out["lf"].cpu().double().numpy().reshape(cfg.p, -1, 2).transpose(0, 2, 1)
    assert np.linalg.norm(lf - GM["lvb_lf_sample"]) <= RTOL * np.linalg.norm(GM["lvb_lf_sample"])
    gth = out["grad_theta"].cpu().double().numpy()
    err = np.linalg.norm(gth - GM["lvb_grad_theta"]) / np.linalg.norm(GM["lvb_grad_theta"])
    print("CUDA vs reference classes (LV batch, learned softplus-theta): worst gradient slice %.2e, d/dtheta %.2e" % (worst, err))
    assert err <= RTOL


def test_train_step_with_a_softplus_posterior_matches_the_host_composition(GM):
    """nma_train_step on the LV batch model: theta = softplus(flow output), Softplus-transformed prior (the two Jacobians
    cancel in prior - log q), against the host autograd module + the oracle's prior on the noise the library drew."""
    from test_step_golden_models import lvb_inputs
    from viforssms_b200.engine import NMAEngine
    from viforssms_b200.theta_flow import ThetaFlow
    cfg, layout, n, params, eps, theta, idx, tf, extra, arrays = lvb_inputs(GM)
    dev = torch.device("cuda")
    eng = NMAEngine(cfg)
    eng.set_series(arrays)
    priors = [(-1.0, np.sqrt(0.1)), (-6.0, np.sqrt(0.1)), (-1.0, np.sqrt(0.1)), (-2.0, np.sqrt(0.1))]
    flow = ThetaFlow(4, 4, 0.0, 1.0, "elu", [np.random.RandomState(k).permutation(4) for k in range(3)], softplus_out=True)
    g = torch.Generator().manual_seed(5)
    # a posterior concentrated near the script's theta*: scale the flow's output layer down and bias it to the prior means
    fp = flow.init_values(g) * 0.1
    blob = torch.cat([params, fp]).to(dev)
    eng.set_theta_flow(flow, priors)
    eng.set_seed(3, 0)
    grad = torch.zeros_like(blob); m = torch.zeros_like(blob); v = torch.zeros_like(blob)
    scal = torch.zeros(8, device=dev)
    before = blob.clone()
    idx_dev = torch.from_numpy(np.asarray(idx)).to(dev)
    eng.train_step(blob, grad, m, v, idx_dev, scal, objective=0, prior_on=True, lr=1e-3, beta1=0.95, clip=1e9)
    torch.cuda.synchronize()
    buf = eng.step_buffers(cfg.p)
    flat = before[n:].clone().requires_grad_(True)
    flow.bind(flat)
    theta_h, lq_h = flow.sample_and_log_prob(buf["z0"])
    assert (theta_h > 0).all() and torch.allclose(theta_h, buf["theta"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(lq_h, buf["logq_theta"], rtol=1e-5, atol=1e-4)
    out = eng.elbo_fwd_bwd(before[:n].contiguous(), buf["eps"], buf["theta"], idx_dev)
    prior = O.lvb_theta_prior(theta_h.double().cpu(), priors).to(dev).float()
    t = out["terms"]
    row = float(cfg.scale) * (t[:, 0] - t[:, 2] + t[:, 1]) + prior - lq_h
    assert torch.allclose(buf["row_elbo"], row.detach(), rtol=1e-4, atol=1e-2)
    mean = torch.tensor([a for a, _ in priors], device=dev); sd = torch.tensor([b for _, b in priors], device=dev, dtype=torch.float32)
    u = theta_h + torch.log(-torch.expm1(-theta_h))
    prior_h = (-0.5 * ((u - mean) / sd) ** 2 - torch.log(sd) - torch.log(-torch.expm1(-theta_h))).sum(1)
    host_loss = (out["grad_theta"] * theta_h).sum() - (prior_h - lq_h).sum()
    host_loss.backward()
    tail = grad[n:]
    err = (tail - flat.grad).norm().item() / flat.grad.norm().item()
    print("softplus posterior through nma_train_step: flow-variable gradient rel err %.2e" % err)
    assert err < 1e-4
    assert torch.allclose(grad[:n], out["grad_params"], rtol=1e-4, atol=1e-5 * out["grad_params"].abs().max().item())


def test_lv_batch_facade_pretrains_trains_and_exports_paths(tmp_path):
    import lotka_volterra_partial_batch as mod
    import lotka_volterra_partial_batch_fix_theta as fix
    np.random.seed(3)
    B, p_val = 31, 3
    obs = fix.simulate(p_val, T=(B - 1) * 0.2, dt=0.2).astype(np.float32)
    assert obs.shape == (2, p_val * B)
    flow = mod.ThetaFlow(4, 4, 0.0, 1.0, "elu", softplus_out=True)
    priors = [(-1.0, np.sqrt(0.1)), (-6.0, np.sqrt(0.1)), (-1.0, np.sqrt(0.1)), (-2.0, np.sqrt(0.1))]
    m = mod.VI_SSM(obs, np.ones_like(obs), np.zeros_like(obs), np.array([91., 99.], np.float32), np.array([1., 1.], np.float32),
                   flow, priors, 0.2, (B - 1) * 0.2, p_val, 20, B, [50] * 5, B, 3, 10, learn_rate=1e-3, pre_train=True)
    m.build_flow()
    for _ in range(3):
        assert m._iteration(m._draw(), pre_train=True) in (True, False)
    for _ in range(4):                                       # eager, capture + replay, replay, replay
        m._iteration(m._draw(), pre_train=False)
    torch.cuda.synchronize()
    sc = m.scalars
    assert set(sc) == {"loss/ELBO", "loss/SDE_log_prob", "loss/theta_log_prob", "loss/obs_log_prob", "loss/path_log_prob",
                       "optimize/global_norm"}
    assert all(np.isfinite(float(v)) for v in sc.values()), sc
    assert (m._theta_last > 0).all()                         # theta lives on the softplus scale
    paths = m.save_paths(str(tmp_path / "paths.txt"))
    assert paths.shape == (p_val, 2, p_val * B) and np.loadtxt(tmp_path / "paths.txt").shape == (p_val, 2 * p_val * B)
    assert (paths >= 1.0).all() and np.isfinite(paths).all()    # states are 1 + softplus(.)
