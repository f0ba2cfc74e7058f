"""Lotka-Volterra NMA model with a learned theta and p_val windows per iteration - drop-in for the reference's
lotka_volterra_partial_batch.py (the file BASELINE.json configs[1] names).

`from lotka_volterra_partial_batch import VI_SSM` gives the class with the reference's constructor and methods (:190-675);
`python lotka_volterra_partial_batch.py` runs what the bottom of the reference script runs (:677-764): load
dat/our_files/LV_{obs_partial,obs_binary,time_till}_test.txt, keep the first p_val = 3 series of 151 steps, build the theta
posterior (4 inverse-MAF layers, base N(0, 1), elu, 3 permutations from numpy's global stream, a final Softplus), build the
model, write posterior paths, train for 3000 epochs.  The TensorFlow graph is replaced by the B200 library
(viforssms_b200/vi_ssm_models.py, LVB_VI_SSM).

The reference ships none of dat/our_files/; `--generate [N_SERIES]` writes synthetic series in that layout with the sibling
script's simulator (Euler-Maruyama of the same SDE).
"""
import os
import sys

import numpy as np

from viforssms_b200.theta_flow import ThetaFlow
from viforssms_b200.vi_ssm_models import LVB_VI_SSM as VI_SSM

NP_DTYPE = np.float32
np.random.seed(1)

__all__ = ["VI_SSM", "main", "generate", "ThetaFlow", "NP_DTYPE"]
DAT = 'dat/our_files'


def generate(n_series=4, T=30, dt=0.2, seed=1, dat_dir=DAT):
    from lotka_volterra_partial_batch_fix_theta import simulate
    obs = simulate(n_series, T, dt, seed)
    os.makedirs(dat_dir, exist_ok=True)
    np.savetxt(os.path.join(dat_dir, 'LV_obs_partial_test.txt'), obs)
    np.savetxt(os.path.join(dat_dir, 'LV_obs_binary_test.txt'), np.ones_like(obs))
    np.savetxt(os.path.join(dat_dir, 'LV_time_till_test.txt'), np.zeros_like(obs))
    return obs


def main(p_val=3, kernel_len=20, dt=0.2, T=30, batch_dims=151, network_dims=(50,) * 5, no_flows=3, feat_window=10,
         num_epochs=3000):
    """:677-764."""
    target_dims = int(np.int32(T / dt)) + 1
    priors = [(-1.0, np.sqrt(0.1)), (-6.0, np.sqrt(0.1)), (-1.0, np.sqrt(0.1)), (-2.0, np.sqrt(0.1))]      # :689-690
    x0_mean = np.array([91., 99.], dtype=NP_DTYPE)
    x0_std = np.array([1., 1.], dtype=NP_DTYPE)
    obs = np.loadtxt(os.path.join(DAT, 'LV_obs_partial_test.txt'), NP_DTYPE)
    obs[obs == -1] = np.log(1 + np.exp(-2)) + 1.0              # f(x) = 1 + softplus(x - 1) for obs = -1 (:707-709)
    obs_bin = np.loadtxt(os.path.join(DAT, 'LV_obs_binary_test.txt'), NP_DTYPE)
    time_till = np.loadtxt(os.path.join(DAT, 'LV_time_till_test.txt'), NP_DTYPE)
    n = p_val * batch_dims
    obs, obs_bin, time_till = obs[:, :n], obs_bin[:, :n], time_till[:, :n]                                   # :717-719
    theta_dist = ThetaFlow(len(priors), 4, base_loc=0., base_scale=1., activation="elu", softplus_out=True)  # :731-747
    var_model = VI_SSM(obs, obs_bin, time_till, x0_mean, x0_std, theta_dist, priors, dt, T, p_val, kernel_len, batch_dims,
                       list(network_dims), target_dims, no_flows, feat_window, learn_rate=1e-3, pre_train=True)
    var_model.build_flow()
    os.makedirs('locally_variant', exist_ok=True)
    var_model.save_paths('locally_variant/LV_obs_paths_series.txt')
    var_model.train(tensorboard_path='locally_variant/train/', save_path='model_saves/LV_model_series_%d_3.ckpt' % batch_dims,
                    series_idx=None, num_epochs=num_epochs)
    return var_model


if __name__ == "__main__":
    if "--generate" in sys.argv:
        k = sys.argv.index("--generate")
        generate(int(sys.argv[k + 1]) if len(sys.argv) > k + 1 else 4)
    else:
        main(num_epochs=int(os.environ.get("NMA_MAX_STEPS", "3000")))
