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
