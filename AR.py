"""AR(1) NMA model — drop-in for the reference's AR.py (classes VI_SSM, main(); AR.py:113-403).

`from AR import *` gives `VI_SSM` and `main` with the reference's signatures.  The TensorFlow graph of the
reference is replaced by the B200 library (viforssms_b200): see viforssms_b200/vi_ssm.py.
"""
import os

import numpy as np

from viforssms_b200.theta_flow import ThetaFlow
from viforssms_b200.vi_ssm import VI_SSM

NP_DTYPE = np.float32
dat_dir = os.getcwd()
np.random.seed(1)                      # AR.py:18

__all__ = ["VI_SSM", "main", "ThetaFlow", "NP_DTYPE", "dat_dir"]


def main(p, kernel_len, T, batch_dims, network_dims, no_flows, priors, feat_window, x0, obs_std, learn_rate=1e-3,
         grad_clip=2.5e8, early_stopping=None):
    """AR.py:364-403: load dat/AR_*.txt, build the theta posterior (5 inverse-MAF layers with 4 random
    permutations drawn from numpy's global stream, base Normal(1.5, 0.5)), build the model, train."""
    obs = np.loadtxt(os.path.join(dat_dir, "dat", "AR_obs_partial.txt"), NP_DTYPE)
    obs_bin = np.loadtxt(os.path.join(dat_dir, "dat", "AR_obs_binary.txt"), NP_DTYPE)
    time_till = np.loadtxt(os.path.join(dat_dir, "dat", "AR_time_till.txt"), NP_DTYPE)

    num_bijectors = 5
    theta_dist = ThetaFlow(len(priors), num_bijectors, base_loc=1.5, base_scale=0.5, activation="elu")
    if early_stopping is None:
        early_stopping = float(os.environ.get("NMA_MAX_STEPS", "1e99"))
    var_model = VI_SSM(obs, obs_std, x0, theta_dist, priors, T, p, kernel_len, batch_dims, network_dims, no_flows,
                       feat_window, obs_bin, time_till, pre_train=True, early_stopping=early_stopping,
                       learn_rate=learn_rate, grad_clip=grad_clip)
    var_model.build_flow()
    var_model.train(tensorboard_path=dat_dir + "/train/", save_path=dat_dir + "/model_saves/AR_save.ckpt")
    return var_model


if __name__ == "__main__":
    main(p=50, kernel_len=50, T=5000, batch_dims=50, network_dims=[50] * 3, no_flows=3, priors=[(0., 10.0)] * 3,
         feat_window=10, x0=10.0, obs_std=1.0)
