"""Lotka-Volterra NMA model with fixed parameters - drop-in for the reference's
lotka_volterra_partial_batch_fix_theta.py.

`from lotka_volterra_partial_batch_fix_theta import VI_SSM` gives the class with the reference's constructor and
methods (:181-613); `python lotka_volterra_partial_batch_fix_theta.py` runs what the bottom of the reference script
runs (:616-709): for each of the series concatenated in dat/our_files/fix_theta/LV_*_dense_test.txt, build a fresh
model (p_val = 1, kernel_len 20, batch_dims 151, 3 flows, 5 x 50 network), write posterior paths, pre-train and
train.  The TensorFlow graph is replaced by the B200 library (viforssms_b200/vi_ssm_models.py, LV_VI_SSM).

The reference ships none of dat/our_files/; `--generate [N_SERIES]` writes synthetic series in that layout with the
script's own SDE (Euler-Maruyama of :274-290 at theta = softplus([-1, -6, -1]), x0 ~ N([91, 99], 1), dt = 0.2, 151
steps per series, observations 1 + softplus(N(x, (theta3 x)^2) - 1), all observed).
"""
import os
import sys

import numpy as np

from viforssms_b200.vi_ssm_models import LV_VI_SSM as VI_SSM

NP_DTYPE = np.float32
np.random.seed(1)

__all__ = ["VI_SSM", "main", "generate", "simulate", "NP_DTYPE"]
DAT = 'dat/our_files/fix_theta'


def softplus_np_(x):
    return np.log(1 + np.exp(x))


def simulate(n_series=4, T=30, dt=0.2, seed=1):
    """Euler-Maruyama of the script's SDE (:274-290), observations 1 + softplus(N(x, (theta3 x)^2) - 1); returns the
    [2, n_series * 151] observation matrix of the dense layout."""
    rs = np.random.RandomState(seed)
    th = softplus_np_(np.array([-1.0, -6.0, -1.0, -2.0]))
    n = int(np.int32(T / dt)) + 1
    obs = np.empty((2, n_series * n))
    for s in range(n_series):
        x = rs.normal([91.0, 99.0], 1.0)
        for t in range(n):
            y = rs.normal(x, th[3] * x)
            obs[:, s * n + t] = 1.0 + softplus_np_(y - 1.0)
            a = np.array([th[0] * x[0] - th[1] * x[0] * x[1], th[1] * x[0] * x[1] - th[2] * x[1]])
            ca = np.sqrt(th[0] * x[0] + th[1] * x[0] * x[1])
            cb = -th[1] * x[0] * x[1] / ca
            cc = np.sqrt(th[1] * x[0] * x[1] + th[2] * x[1] - cb ** 2)
            z = rs.standard_normal(2)
            x = np.maximum(x + dt * a + np.sqrt(dt) * np.array([ca * z[0], cb * z[0] + cc * z[1]]), 1.5)
    return obs


def generate(n_series=4, T=30, dt=0.2, seed=1, dat_dir=DAT):
    obs = simulate(n_series, T, dt, seed)
    os.makedirs(dat_dir, exist_ok=True)
    np.savetxt(os.path.join(dat_dir, 'LV_obs_partial_dense_test.txt'), obs)
    np.savetxt(os.path.join(dat_dir, 'LV_obs_binary_dense_test.txt'), np.ones_like(obs))
    np.savetxt(os.path.join(dat_dir, 'LV_time_till_dense_test.txt'), np.zeros_like(obs))
    return obs


def main(n_series=150, num_epochs=3000, pre_train_epochs=1000, p_val=1, kernel_len=20, dt=0.2, T=30, batch_dims=151,
         network_dims=(50,) * 5, no_flows=3, feat_window=10):
    """:616-709."""
    target_dims = int(np.int32(T / dt)) + 1
    priors = softplus_np_(np.array([-1.0, -6.0, -1.0, -2.0]))
    x0_mean = np.array([91., 99.], dtype=NP_DTYPE)
    x0_std = np.array([1., 1.], dtype=NP_DTYPE)
    obs_all = np.loadtxt(os.path.join(DAT, 'LV_obs_partial_dense_test.txt'), NP_DTYPE)
    obs_all[obs_all == -1] = np.log(1 + np.exp(-2)) + 1.0          # f(x) = 1 + softplus(x - 1) for obs = -1 (:661-663)
    bin_all = np.loadtxt(os.path.join(DAT, 'LV_obs_binary_dense_test.txt'), NP_DTYPE)
    tt_all = np.loadtxt(os.path.join(DAT, 'LV_time_till_dense_test.txt'), NP_DTYPE)
    os.makedirs('locally_variant/fix_theta', exist_ok=True)
    models = []
    for idx in range(min(n_series, obs_all.shape[1] // batch_dims)):
        print('=' * 50)
        print('Starting series %d' % idx)
        sl = slice(idx * batch_dims, (idx + 1) * batch_dims)
        var_model = VI_SSM(obs_all[:, sl], bin_all[:, sl], tt_all[:, sl], x0_mean, x0_std, priors, dt, T, p_val,
                           kernel_len, batch_dims, list(network_dims), target_dims, no_flows, feat_window,
                           learn_rate=1e-3, pre_train=True)
        var_model.build_flow()
        var_model.save_paths('locally_variant/fix_theta/LV_obs_paths_series_dense_%d.txt' % idx)
        var_model.train(tensorboard_path='locally_variant/fix_theta/train_dense/',
                        save_path='model_saves/fix_theta/LV_model_series_%d_3_dense_%d.ckpt' % (batch_dims, idx),
                        num_epochs=num_epochs, pre_train_epochs=pre_train_epochs, series_idx=idx)
        os.makedirs('dat/our_files', exist_ok=True)
        np.save('dat/our_files/series_%d_learned_lf_sample_dense.npy' % idx, var_model.lf_sample.cpu().numpy())
        models.append(var_model)
    print('All series done')
    return models


if __name__ == "__main__":
    if "--generate" in sys.argv:
        k = sys.argv.index("--generate")
        generate(int(sys.argv[k + 1]) if len(sys.argv) > k + 1 else 4)
    else:
        main(num_epochs=int(os.environ.get("NMA_MAX_STEPS", "3000")))
