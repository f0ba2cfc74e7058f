"""GPU tests of the whole-iteration entry point (nma_train_step, viforssms_b200/csrc/nma_step.cu): in-library Philox
noise (eps == NULL of SURVEY section 8b), the oracle on what it drew, CUDA-graph replay, Adamax on unaligned views."""
import os

import numpy as np
import pytest
import torch

from oracle import nma_oracle as O

pytestmark = [pytest.mark.gpu]


# ---- Philox4x32-10 + Box-Muller in numpy (Salmon et al. 2011): the generator of nma_step.cu ----
def philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [np.asarray(x, dtype=np.uint64) for x in ctr]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF), p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def philox_normals(n, seed, counter, stream_id):
    nb = (n + 3) // 4
    b = np.arange(nb, dtype=np.uint64)
    ctr = [b & np.uint64(0xFFFFFFFF), b >> np.uint64(32), np.full(nb, counter & 0xFFFFFFFF, np.uint64),
           np.full(nb, ((counter >> 32) ^ (stream_id << 24)) & 0xFFFFFFFF, np.uint64)]
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32))
    def bm(a, bb):
        u1 = ((a >> np.uint64(8)).astype(np.float64) + 1.0) / 16777216.0
        u2 = (bb >> np.uint64(8)).astype(np.float64) / 16777216.0
        rad = np.sqrt(-2.0 * np.log(u1))
        return rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)
    n0, n1 = bm(r[0], r[1])
    n2, n3 = bm(r[2], r[3])
    return np.stack([n0, n1, n2, n3], axis=1).reshape(-1)[:n]


def test_philox_known_answer():
    # Random123's published known-answer vector for philox4x32-10: counter = key = 0
    r = philox4x32_10([np.zeros(1, np.uint64)] * 4, (0, 0))
    assert [int(x[0]) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = philox4x32_10([np.full(1, 0xFFFFFFFF, np.uint64)] * 4, (0xFFFFFFFF, 0xFFFFFFFF))
    assert [int(x[0]) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


@pytest.mark.parametrize("n,seed,counter,stream", [(10007, 1, 0, 0), (4096, 0x1234567890, 77, 1), (5, 3, 2 ** 33 + 5, 0)])
def test_library_noise_is_philox_box_muller(n, seed, counter, stream):
    from viforssms_b200.engine import philox_normal
    got = philox_normal(n, seed, counter, stream, loc=0.5, scale=2.0).cpu().double().numpy()
    want = 0.5 + 2.0 * philox_normals(n, seed, counter, stream)
    assert np.abs(got - want).max() < 2e-5
    big = philox_normal(1 << 22, seed, counter, stream).cpu().double().numpy()
    assert abs(big.mean()) < 3e-3 and abs(big.std() - 1.0) < 3e-3
    other = philox_normal(1 << 12, seed, counter + 1, stream).cpu().numpy()
    assert np.abs(other - big[:1 << 12]).max() > 1.0              # a new counter is a new draw


def _ar_setup(p=24, T=5000, seed=3):
    from viforssms_b200 import feed
    from viforssms_b200.config import ar_config, param_layout
    from viforssms_b200.engine import NMAEngine
    from viforssms_b200.theta_flow import ThetaFlow
    from viforssms_b200.trainer import glorot_blob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "dat")
    obs = np.loadtxt(os.path.join(d, "AR_obs_partial.txt"))
    obs_bin = np.loadtxt(os.path.join(d, "AR_obs_binary.txt"))
    tt = np.loadtxt(os.path.join(d, "AR_time_till.txt"))
    cfg = ar_config(p=p, K=8, B=10, F=2, H=1, feat_window=4, T=T)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(seed)
    nma = glorot_blob(cfg, g)
    for i in range(cfg.F):
        off, shape = layout[f"f{i}.feat0.w"]
        nma[off:off + shape[0] * shape[1]].reshape(shape)[cfg.Cf - 3, :] *= 10.0 / T
    flow = ThetaFlow(3, 5, 1.5, 0.5, "elu", [np.random.RandomState(k).permutation(3) for k in range(4)])
    dev = torch.device("cuda", 0)
    blob = torch.cat([nma, flow.init_values(g)]).to(dev)
    eng = NMAEngine(cfg, dev)
    eng.set_series(feed.ar_base_arrays(obs, obs_bin, tt, T, cfg.F, cfg.K, 4))
    priors = [(0.0, 10.0)] * 3
    eng.set_theta_flow(flow, priors)
    idx = torch.from_numpy(feed.sample_indices(T, cfg.B, p, np.random.RandomState(seed))).to(dev)
    pads = O.pad_series_ar(obs, obs_bin, tt, 10.0, T, cfg.F, cfg.K, 4)
    tf64, _, _ = O.gather_feed_ar(pads, idx.cpu().numpy(), cfg.L0, cfg.B)
    tf = torch.from_numpy(tf64.astype(np.float32)).double()
    return cfg, layout, n, flow, blob, eng, idx, tf, priors


def test_train_step_against_the_oracle_on_the_noise_it_drew():
    """nma_train_step's ELBO terms, mean ELBO and NMA gradient against the fp64 oracle fed the eps / theta the library
    drew (read back from the workspace); the draw counter advances by one per call."""
    from viforssms_b200.engine import philox_normal
    cfg, layout, n, flow, blob, eng, idx, tf, priors = _ar_setup()
    dev = blob.device
    eng.set_seed(11, 5)
    grad = torch.zeros_like(blob); m = torch.zeros_like(blob); v = torch.zeros_like(blob)
    scal = torch.zeros(8, device=dev)
    before = blob.clone()
    eng.train_step(blob, grad, m, v, idx, scal, objective=0, prior_on=True, lr=1e-3, beta1=0.95, clip=2.5e8)
    torch.cuda.synchronize()
    buf = eng.step_buffers(cfg.p)
    assert eng.draw_counter() == 6 and scal[7].item() == 6.0
    assert torch.equal(buf["eps"].reshape(-1), philox_normal(cfg.p * cfg.L0, 11, 5, 0, device=dev))
    assert torch.equal(buf["z0"].reshape(-1), philox_normal(cfg.p * 3, 11, 5, 1, 1.5, 0.5, device=dev))
    ref = O.step_reference(cfg, layout, before[:n].cpu().double(), buf["eps"].cpu().double(), buf["theta"].cpu().double(), tf)
    t = buf["terms"].cpu().double()
    assert (t - ref["terms"]).abs().max().item() <= 1e-4 * max(1.0, ref["terms"].abs().max().item())
    gerr = ((grad[:n].cpu().double() - ref["grad_params"]).norm() / ref["grad_params"].norm()).item()
    assert gerr < 1e-4
    # the tail: log prior(theta) - log q(theta) (AR.py:178-185), through the host module
    flat = before[n:].clone().requires_grad_(True)
    flow.bind(flat)
    theta_h, lq_h = flow.sample_and_log_prob(buf["z0"])
    assert torch.allclose(theta_h, buf["theta"], rtol=1e-5, atol=1e-5)
    th = buf["theta"].cpu().double()
    prior = (-0.5 * (th / 10.0) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(10.0)).sum(1)
    rt = ref["terms"]
    row = cfg.scale * (rt[:, 0] - rt[:, 2] + rt[:, 1]) + prior - lq_h.detach().cpu().double()
    assert (buf["row_elbo"].cpu().double() - row).abs().max().item() <= 1e-4 * row.abs().max().item()
    assert abs(scal[0].item() - row.mean().item()) <= 1e-4 * abs(row.mean().item())
    # global norm and the update
    assert abs(scal[5].item() - grad.norm().item()) <= 1e-5 * grad.norm().item()
    w2, _, _ = O.adamax_step(before.cpu(), grad.cpu(), torch.zeros(blob.numel()), torch.zeros(blob.numel()), 1e-3, 0.95,
                             clip=(2.5e8, float(grad.norm())))
    mask = flow.mask_flat()
    w2[n:] *= mask
    assert torch.allclose(blob.cpu(), w2, rtol=0, atol=2e-6)


def test_eps_null_draws_what_nma_philox_normal_reports():
    from viforssms_b200.engine import philox_normal
    cfg, layout, n, flow, blob, eng, idx, tf, priors = _ar_setup()
    dev = blob.device
    eng.set_seed(21, 0)
    theta = (torch.tensor([4.0, 0.5, 1.0]) + 0.1 * torch.randn(cfg.p, 3)).to(dev)
    a = eng.elbo_fwd_bwd(blob[:n].contiguous(), None, theta, idx)
    torch.cuda.synchronize()
    assert eng.draw_counter() == 1
    eps = philox_normal(cfg.p * cfg.L0, 21, 0, 0, device=dev).reshape(cfg.p, cfg.L0)
    b = eng.elbo_fwd_bwd(blob[:n].contiguous(), eps, theta, idx)
    torch.cuda.synchronize()
    assert eng.draw_counter() == 1
    assert torch.equal(a["terms"], b["terms"])
    assert torch.allclose(a["grad_params"], b["grad_params"], rtol=1e-4, atol=1e-6 * b["grad_params"].abs().max().item())
    terms, lf = eng.forward_paths(blob[:n].contiguous(), None, theta, idx)
    assert eng.draw_counter() == 2 and torch.isfinite(terms).all()


def test_graph_replay_equals_eager_iterations(monkeypatch):
    """The AR facade replays its iteration from a CUDA graph; the noise is counter-based, so the graph run and the
    eager run draw the same numbers and must arrive at the same variables (up to atomics' summation order)."""
    import AR as ar_mod
    from viforssms_b200.theta_flow import ThetaFlow
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "dat")
    obs = np.loadtxt(os.path.join(d, "AR_obs_partial.txt"))
    obs_bin = np.loadtxt(os.path.join(d, "AR_obs_binary.txt"))
    tt = np.loadtxt(os.path.join(d, "AR_time_till.txt"))
    blobs, scal = [], []
    for graph in ("0", "1"):
        monkeypatch.setenv("NMA_FACADE_GRAPH", graph)
        np.random.seed(1)
        flow = ThetaFlow(3, 5, 1.5, 0.5, "elu")
        m = ar_mod.VI_SSM(obs, 1.0, 10.0, flow, [(0.0, 10.0)] * 3, 5000, 20, 10, 10, [50, 50, 50], 2, 4, obs_bin, tt, seed=2)
        m.build_flow()
        for it in range(6):
            m._iteration(m._draw(False), pre_train=it < 3)
        torch.cuda.synchronize()
        blobs.append(m.blob.clone()); scal.append(m.scalars)
        assert m.eng.draw_counter() == 6
        assert (len(m._graphs) == 2) == (graph == "1")
    # same draws, same arithmetic; the weight gradient is summed with atomics, so entries whose gradient is rounding noise may
    # take an Adamax step of the other sign (|step| <= lr): a handful of entries, bounded by 2 * lr per iteration
    diff = (blobs[0] - blobs[1]).abs()
    assert (diff > 2e-5).float().mean().item() < 2e-3 and diff.max().item() <= 6 * 2e-3
    for k in scal[0]:
        assert abs(scal[0][k] - scal[1][k]) <= 1e-4 * max(1.0, abs(scal[0][k])), k


def test_second_stream_train_step_equals_one_stream(monkeypatch):
    """nma_train_step at the scripts' row counts runs flow 0's conv / feature backward on the handle's second stream next to
    the theta-bias MLP backward, the prior terms and the theta posterior's backward, and joins right before the optimiser
    (nma_api.cu, nma_step.cu).  NMA_NO_AUX_STREAM=1 keeps one stream: same draws, same variables up to the atomics' order -
    eagerly and from the captured graph."""
    import AR as ar_mod
    from viforssms_b200.theta_flow import ThetaFlow
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "dat")
    obs = np.loadtxt(os.path.join(d, "AR_obs_partial.txt"))
    obs_bin = np.loadtxt(os.path.join(d, "AR_obs_binary.txt"))
    tt = np.loadtxt(os.path.join(d, "AR_time_till.txt"))
    blobs, scal = [], []
    for no_aux, graph in (("1", "0"), (None, "0"), (None, "1")):
        monkeypatch.setenv("NMA_FACADE_GRAPH", graph)
        if no_aux:
            monkeypatch.setenv("NMA_NO_AUX_STREAM", no_aux)
        else:
            monkeypatch.delenv("NMA_NO_AUX_STREAM", raising=False)
        np.random.seed(1)
        flow = ThetaFlow(3, 5, 1.5, 0.5, "elu")
        m = ar_mod.VI_SSM(obs, 1.0, 10.0, flow, [(0.0, 10.0)] * 3, 5000, 20, 10, 10, [50, 50, 50], 2, 4, obs_bin, tt, seed=2)
        m.build_flow()
        for it in range(6):
            m._iteration(m._draw(False), pre_train=it < 3)
        torch.cuda.synchronize()
        blobs.append(m.blob.clone()); scal.append(m.scalars)
    for other in (1, 2):
        diff = (blobs[0] - blobs[other]).abs()
        assert (diff > 2e-5).float().mean().item() < 2e-3 and diff.max().item() <= 6 * 2e-3
        for k in scal[0]:
            assert abs(scal[0][k] - scal[other][k]) <= 1e-4 * max(1.0, abs(scal[0][k])), k


@pytest.mark.parametrize("offset", [1, 2, 3, 5])
def test_adamax_on_unaligned_tail_views(offset):
    """The second pre-train optimiser updates views into the tail of the blob (vi_ssm_models.py): any 4-byte offset."""
    from viforssms_b200.config import ar_config
    from viforssms_b200.engine import NMAEngine
    dev = torch.device("cuda", 0)
    eng = NMAEngine(ar_config(p=2, K=4, B=4, F=1, H=1, feat_window=2, T=100), dev)
    g = torch.Generator().manual_seed(offset)
    n = 1003
    w = torch.randn(n + 8, generator=g); gr = torch.randn(n + 8, generator=g)
    m = torch.rand(n + 8, generator=g); v = torch.randn(n + 8, generator=g)
    wd, gd, md, vd = (t.to(dev) for t in (w, gr, m, v))
    sl = slice(offset, offset + n)
    norm = eng.adamax_step(wd[sl], gd[sl], md[sl], vd[sl], 1e-2, 0.9, clip=3.0)
    torch.cuda.synchronize()
    gn = float(gr[sl].double().norm())
    assert abs(norm.item() - gn) <= 1e-5 * gn
    w2, m2, v2 = O.adamax_step(w[sl].clone(), gr[sl], m[sl].clone(), v[sl].clone(), 1e-2, 0.9, clip=(3.0, gn))
    assert torch.allclose(wd[sl].cpu(), w2, rtol=1e-5, atol=1e-6)
    # nothing outside the view moved
    assert torch.equal(wd[:offset].cpu(), w[:offset]) and torch.equal(wd[offset + n:].cpu(), w[offset + n:])


@pytest.mark.parametrize("n", [1, 17, 4096, 4097, 4096 * 70 + 123, 3_000_001])
def test_lookback_scan_matches_the_sequential_recursion(n):
    """A12 (AR_dat_gen.py:11-15) as a single-pass decoupled look-back scan: one tile, a ragged tile, more than 32 tiles
    (several look-back windows), ~730 tiles; the per-element form as well."""
    from viforssms_b200.engine import scan_affine, scan_ar1
    dev = torch.device("cuda", 0)
    rs = np.random.RandomState(n % 1000)
    z = rs.standard_normal(n)
    a, b, c, x0 = 0.5, 5.0, 3.0, 10.0
    want = np.empty(n + 1); want[0] = x0
    A = rs.uniform(0.5, 1.05, size=n); D = rs.standard_normal(n)
    want2 = np.empty(n + 1); want2[0] = -2.0
    x, y = x0, -2.0
    for i in range(n):
        x = a * x + b + c * z[i]; want[i + 1] = x
        y = A[i] * y + D[i]; want2[i + 1] = y
    for rep in range(2):                      # twice: the scratch (ticket, flags) is reset by every call
        got = scan_ar1(torch.from_numpy(z).to(dev), x0, a, b, c).cpu().numpy()
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    got2 = scan_affine(torch.from_numpy(A).to(dev), torch.from_numpy(D).to(dev), -2.0).cpu().numpy()
    assert np.abs(got2 - want2).max() <= 1e-11 * max(1.0, np.abs(want2).max())


def test_sv_generator_on_the_device_scans_matches_the_host_loop():
    import SV_dense as mod
    host = mod.simulate(1809)
    devs = mod.simulate(1809, device=torch.device("cuda", 0))
    assert np.allclose(devs, host, rtol=1e-9, atol=0)
