"""Lotka-Volterra NMA model with a learned theta posterior - drop-in for the reference's lotka_volterra_partial.py.

`from lotka_volterra_partial import VI_SSM` gives the class with the reference's constructor and methods
(lotka_volterra_partial.py:160-463); `python lotka_volterra_partial.py` runs what the bottom of the reference script
runs (:465-524): load dat/LV_{obs_partial,obs_binary,time_till}.txt (the reference ships these three files; copy them
into dat/), build the theta posterior (4 inverse-MAF layers, base N(0, 1), elu, 3 permutations from numpy's global
stream), build the model, save posterior paths under locally_variant/, train.  The TensorFlow graph is replaced by the
B200 library (viforssms_b200/vi_ssm_models.py, LVR_VI_SSM).

STATUS: the model's oracle is pinned to the reference script's own classes (tests/test_step_golden_models.py) and the
kernel's hand-derived gradients to the oracle (tests/test_lvr_formulas.py), both on the CPU; the ELBO kernel branch has
not been run on a B200 yet, so the library refuses the model unless NMA_UNVERIFIED=1 is set
(tests/test_gpu_unverified.py is the parity test to run first).
"""
import os
import sys

import numpy as np

from viforssms_b200.theta_flow import ThetaFlow
from viforssms_b200.vi_ssm_models import LVR_VI_SSM as VI_SSM

NP_DTYPE = np.float32
np.random.seed(1)                      # lotka_volterra_partial.py:19

__all__ = ["VI_SSM", "main", "ThetaFlow", "NP_DTYPE"]

PRIORS = ((np.log(4.428 / 10), 1e-4), (np.log(0.029 / 10), 1e-4), (np.log(2.957 / 10), 1e-4))      # :476


def main(train=True, p=50, kernel_len=20, dt=0.1, T=50., batch_dims=50, network_dims=(50,) * 5, no_flows=3,
         priors=PRIORS, feat_window=10, x0=(100., 100.), learn_rate=1e-3, early_stopping=None):
    """lotka_volterra_partial.py:465-524."""
    obs = np.loadtxt('dat/LV_obs_partial.txt', NP_DTYPE)
    obs_bin = np.loadtxt('dat/LV_obs_binary.txt', NP_DTYPE)
    time_till = np.loadtxt('dat/LV_time_till.txt', NP_DTYPE)
    target_dims = int(np.int32(T / dt))
    theta_dist = ThetaFlow(len(priors), 4, base_loc=0., base_scale=1., activation="elu")
    if early_stopping is None:
        early_stopping = float(os.environ.get("NMA_MAX_STEPS", "1e99"))
    var_model = VI_SSM(obs, obs_bin, time_till, np.array(x0), theta_dist, list(priors), dt, T, p, kernel_len,
                       batch_dims, list(network_dims), target_dims, no_flows, feat_window, learn_rate=learn_rate,
                       pre_train=True, early_stopping=early_stopping)
    var_model.build_flow()
    os.makedirs('locally_variant', exist_ok=True)
    var_model.save_paths('locally_variant/LV_obs_paths.txt')
    if train:
        var_model.train(tensorboard_path='locally_variant/train/',
                        save_path='model_saves/LV_model_%i_3.ckpt' % batch_dims)
    return var_model


if __name__ == "__main__":
    main(train="--no-train" not in sys.argv)
