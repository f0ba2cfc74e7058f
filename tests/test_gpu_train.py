"""GPU tests of the training surface: VI_SSM (the reference's class, AR.py:113-362) and the device-resident stepper."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _series(T, seed=3):
    import AR_dat_gen
    np.random.seed(seed)
    fill, binary, till = AR_dat_gen.simulate(T, 1, 10.0, np.array([5.0, 0.5, 3.0]), 1.0)
    return fill.astype(np.float32), binary.astype(np.float32), till.astype(np.float32)


def _model(T=1000, p=32, K=10, B=20, F=2, fw=3, pre_train=False, early_stopping=30, seed=1):
    from viforssms_b200.theta_flow import ThetaFlow
    from viforssms_b200.vi_ssm import VI_SSM
    obs, obs_bin, tt = _series(T)
    np.random.seed(seed)
    theta_dist = ThetaFlow(3, 5, 1.5, 0.5, "elu")
    m = VI_SSM(obs, 1.0, 10.0, theta_dist, [(0., 10.0)] * 3, T, p, K, B, [50, 50, 50], F, fw, obs_bin, tt,
               pre_train=pre_train, early_stopping=early_stopping, learn_rate=1e-3, grad_clip=2.5e8, seed=seed)
    m.build_flow()
    return m


def test_vi_ssm_trains_saves_restores_and_exports_paths(tmp_path):
    m = _model()
    w0 = m.blob.clone()
    m.train(str(tmp_path / "train"), str(tmp_path / "model_saves" / "AR_save.ckpt"))
    assert torch.isfinite(m.blob).all() and not torch.equal(m.blob, w0)
    assert all(np.isfinite(float(v)) for v in m.scalars.values())
    assert set(m.scalars) == {"loss/ELBO", "loss/SDE_log_prob", "loss/theta_log_prob", "loss/obs_log_prob",
                              "loss/path_log_prob", "optimize/global_norm"}
    assert os.path.exists(tmp_path / "model_saves" / "AR_save.ckpt")
    # checkpoint round trip: weights and both optimisers' slots
    m.save(str(tmp_path / "ck.pt"))
    w1 = m.blob.clone()
    m.blob.zero_()
    m.load(str(tmp_path / "ck.pt"))
    assert torch.equal(m.blob, w1) and m.pre_train is False
    # posterior paths: [p, T] like the reference's np.savetxt dump
    paths = m.save_paths(str(tmp_path / "paths.txt"))
    assert paths.shape == (m.p, 1000) and np.isfinite(paths).all()
    assert np.loadtxt(tmp_path / "paths.txt").shape == (m.p, 1000)


def test_vi_ssm_pretrain_then_elbo_improves():
    """-obs_loss pre-training pulls the path onto the observations; the main objective then raises the ELBO."""
    m = _model(T=1000, p=64, pre_train=False, early_stopping=0)
    m.train("/tmp/nma_tb_test", "/tmp/nma_tb_test/save.ckpt")
    first = float(m.scalars["loss/ELBO"])
    obs_before = float(m.scalars["loss/obs_log_prob"])
    rep = bool(m.batch_dims * m.p >= m.T)
    for _ in range(150):
        m._iteration(m._draw(rep), pre_train=True)
    m._iteration(m._draw(rep), pre_train=False)
    assert float(m.scalars["loss/obs_log_prob"]) > obs_before
    for _ in range(300):
        m._iteration(m._draw(rep), pre_train=False)
    last = np.mean([float(m.scalars["loss/ELBO"])])
    assert np.isfinite(last) and last > first


def test_stepper_resident_and_e2e_agree_on_shapes_and_progress():
    from viforssms_b200.trainer import ARStepper
    st = ARStepper(T=200000, rows=256, K=50, B=50, F=3, H=1, fw=10, device=torch.device("cuda", 0))
    try:
        assert st.eng.tensor_cores
        e0 = float(st.step_resident().item())
        for _ in range(5):
            st.step_resident()
        e1 = st.step_e2e()
        assert np.isfinite(e0) and np.isfinite(e1)
        assert torch.isfinite(st.blob).all()
        assert st.h2d_bytes == 256 * 8
    finally:
        st.close()


def test_stepper_cuda_graph_step():
    """The whole iteration captured as one CUDA graph: replays keep training and stay finite."""
    from viforssms_b200.trainer import ARStepper
    st = ARStepper(T=100000, rows=64, device=torch.device("cuda", 0))
    try:
        st.capture()
        assert st.launches_per_step and st.launches_per_step > 10
        w0 = st.blob.clone()
        vals = [float(st.step_resident().item()) for _ in range(20)]
        vals.append(st.step_e2e())
        assert all(np.isfinite(v) for v in vals)
        assert torch.isfinite(st.blob).all() and not torch.equal(st.blob, w0)
    finally:
        st.close()


def test_stepper_sharded_series_matches_unsharded_windows():
    """A rank's local arrays (rank 1 of 2, halos taken from the host series) give the same gathered windows as
    the unsharded engine."""
    from viforssms_b200 import feed
    from viforssms_b200.config import ar_config
    from viforssms_b200.engine import NMAEngine
    from viforssms_b200.trainer import ARStepper
    T = 4000
    obs, obs_bin, tt = _series(T, seed=9)
    full = NMAEngine(ar_config(p=8, T=T))
    full.set_series(feed.ar_base_arrays(obs, obs_bin, tt, T, 3, 50, 10))
    st = ARStepper(T=T, rows=8, device=torch.device("cuda", 0), rank=1, world=2, series=(obs, obs_bin, tt))
    try:
        assert (st.t0, st.t1) == (2000, 4000)
        idx = np.array([2000, 2050, 2500, 3000, 3500, 3900, 3950, 2100])
        want, _, _ = full.gather(idx)
        got, _, _ = st.eng.gather(idx - st.t0)
        assert torch.equal(got, want)
    finally:
        st.close()


# ---------------------------------------------------------------------------------------------------
# FitzHugh-Nagumo and stochastic-volatility facades (fitz_nag_NVP.py / SV_dense.py surfaces)
# ---------------------------------------------------------------------------------------------------

def _fhn_model(N=2000, p=16, K=10, B=20, F=2, fw=3, pre_train=False, early_stopping=5):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import synth
    from viforssms_b200.theta_flow import ThetaFlow
    from fitz_nag_NVP import VI_SSM
    obs, obs_bin, tt = synth.fhn_inputs(N)
    np.random.seed(1)
    theta_dist = ThetaFlow(5, 4, 0.0, 1.0, "elu")
    m = VI_SSM(obs, obs_bin, tt, np.array([2.0, 3.0]), theta_dist, [(0., 10.)] * 5, 0.1, N * 0.1, p, K, B, [50] * 5, N,
               F, fw, learn_rate=1e-4, pre_train=pre_train, early_stopping=early_stopping)
    m.build_flow()
    return m


def test_fhn_facade_pretrains_trains_and_exports_paths(tmp_path):
    m = _fhn_model()
    w0 = m.blob.clone()
    th0 = m.blob[m.n_nma:].clone()
    for _ in range(3):                                   # the two pre-train optimisers in one step
        assert m._iteration(m._draw(), pre_train=True)
    assert not torch.equal(m.blob[:m.n_nma], w0[:m.n_nma]) and not torch.equal(m.blob[m.n_nma:], th0)
    assert all(float(s[0].abs().sum()) > 0 for s in (m.slots["pre_path"], m.slots["pre_theta"]))
    assert float(m.slots["pre_theta"][0][:m.n_nma].abs().sum()) == 0.0     # t2 touches the theta flow only
    m.train(str(tmp_path / "train"), str(tmp_path / "model_saves" / "fhn.ckpt"))
    assert torch.isfinite(m.blob).all()
    assert set(m.scalars) == {"loss/ELBO", "loss/SDE_log_prob", "loss/theta_log_prob", "loss/obs_log_prob",
                              "loss/path_log_prob", "optimize/global_norm"}
    assert all(np.isfinite(float(v)) for v in m.scalars.values())
    paths = m.save_paths(str(tmp_path / "paths.txt"))
    assert paths.shape == (m.p, 2, 2000) and np.isfinite(paths).all()
    assert np.loadtxt(tmp_path / "paths.txt").shape == (m.p, 4000)
    m.save(str(tmp_path / "ck.pt"))
    w1 = m.blob.clone()
    m.blob.zero_()
    m.load(str(tmp_path / "ck.pt"))
    assert torch.equal(m.blob, w1)


def test_sv_facade_pretrains_trains_and_exports_paths(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import synth
    from viforssms_b200.theta_flow import ThetaFlow
    from SV_dense import VI_SSM
    obs = synth.sv_prices(686)[100:]                     # 586 prices: T = 585 = 45 * 13, every window fits the series
    T = obs.shape[0] - 1                                 # (as 1508 = 29 * 52 does in the script, SV_dense.py:404-413)
    np.random.seed(1)
    theta_dist = ThetaFlow(4, 5, 0.0, 1.0, "relu")
    m = VI_SSM(obs, -8.5, theta_dist, [(0., 10.0)] * 4, 1.0, T, 24, 10, 13, [50] * 5, T, 3, 2, learn_rate=1e-4,
               pre_train=False, early_stopping=5)
    m.build_flow()
    assert m.eng.tensor_cores
    for _ in range(3):
        m._iteration(m._draw(), pre_train=True)
    m.train(str(tmp_path / "train"), str(tmp_path / "model_saves" / "sv.ckpt"))
    assert torch.isfinite(m.blob).all()
    assert set(m.scalars) == {"loss/ELBO", "loss/SDE_log_prob", "loss/theta_log_prob", "loss/path_log_prob",
                              "optimize/global_norm"}
    paths = m.save_paths(str(tmp_path / "paths.txt"))
    n = len(np.arange(0, T, 13)) * 13
    assert paths.shape == (m.p, 2, n) and np.isfinite(paths).all()
    # component 0 is the observed price itself, component 1 the latent log-volatility
    assert np.allclose(paths[0, 0, :T - 1], obs[1:T], rtol=1e-6)


def test_lv_facade_script_pretrains_trains_and_exports_paths(tmp_path, monkeypatch):
    """lotka_volterra_partial_batch_fix_theta.py surface at the script's own shape (p_val = 1, 151 steps, 364-slot
    window): generate two synthetic series, run the script body for a few epochs."""
    monkeypatch.chdir(tmp_path)
    import lotka_volterra_partial_batch_fix_theta as lv
    obs = lv.generate(n_series=2)
    assert obs.shape == (2, 302) and (obs > 1.0).all()
    models = lv.main(n_series=2, num_epochs=12, pre_train_epochs=6)
    assert len(models) == 2
    for m in models:
        assert not m.pre_train                                   # 6 finite pre-train steps, then 6 ELBO steps
        assert torch.isfinite(m.blob).all()
        assert all(np.isfinite(float(v)) for v in m.scalars.values())
        assert m.lf_sample.shape == (1, 2, 152)
        assert float(m.lf_sample[0, 0, 0]) == 91.0 and float(m.lf_sample[0, 1, 0]) == 99.0     # pinned by mask/shift
        assert (m.lf_sample[:, :, 1:] > 1.0).all()               # 1 + softplus(.)
    assert np.loadtxt(tmp_path / "locally_variant/fix_theta/LV_obs_paths_series_dense_1.txt").shape == (302,)
    assert os.path.exists(tmp_path / "model_saves/fix_theta/LV_model_series_151_3_dense_0.ckpt")
