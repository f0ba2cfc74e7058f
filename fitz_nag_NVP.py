"""FitzHugh-Nagumo NMA model with the RealNVP-style coupling flow - drop-in for the reference's fitz_nag_NVP.py.

`from fitz_nag_NVP import VI_SSM` gives the class with the reference's constructor and methods
(fitz_nag_NVP.py:158-448); `python fitz_nag_NVP.py` runs what the bottom of the reference script runs
(:451-523): load dat/fitz_nag_{obs_partial,obs_binary,time_till}.txt, build the theta posterior (4 inverse-MAF
layers, base N(0, 1), elu, 3 permutations from numpy's global stream), build the model, save posterior paths and
theta samples under locally_variant/, and - with --train - train.  The TensorFlow graph is replaced by the B200
library (viforssms_b200/vi_ssm_models.py).

The reference repository does not ship the FHN series; `python fitz_nag_NVP.py --generate [N]` writes a synthetic
one with the script's own SDE (Euler-Maruyama of fitz_nag_NVP.py:243-255 at theta*, both components observed every
10th step with sd 0.1), in the three-file layout the script reads.
"""
import os
import sys

import numpy as np

from viforssms_b200.theta_flow import ThetaFlow
from viforssms_b200.vi_ssm_models import FHN_VI_SSM as VI_SSM

NP_DTYPE = np.float32
np.random.seed(1)                      # fitz_nag_NVP.py:21

__all__ = ["VI_SSM", "main", "generate", "simulate", "ThetaFlow", "NP_DTYPE"]

THETA_STAR = (np.log(2.), 1., 1.5, np.log(.5), np.log(.3))      # fitz_nag_NVP.py:292


def simulate(target_dims=1000000, dt=0.1, x0=(2., 3.), obs_every=10, obs_std=0.1, seed=1):
    """Euler-Maruyama of dX = alpha dt + sqrt(beta) dW with the drift / diffusion of fitz_nag_NVP.py:243-255; returns
    (obs, obs_bin, time_till), each [2, target_dims], in the layout the script reads."""
    rs = np.random.RandomState(seed)
    th = THETA_STAR
    x = np.empty((2, target_dims + 1))
    x[:, 0] = x0
    s1, s2 = np.sqrt(dt * np.exp(th[3])), np.sqrt(dt * np.exp(th[4]))
    z = rs.standard_normal((2, target_dims))
    for t in range(target_dims):
        a, b = x[0, t], x[1, t]
        x[0, t + 1] = a + dt * np.exp(th[0]) * (a - a ** 3 - b + th[1]) + s1 * z[0, t]
        x[1, t + 1] = b + dt * (th[2] * a - b + 1.4) + s2 * z[1, t]
    lat = x[:, 1:]
    noisy = lat + obs_std * rs.standard_normal(lat.shape)
    obs_bin = np.zeros_like(lat)
    obs_bin[:, obs_every - 1::obs_every] = 1.0
    nxt = np.minimum(((np.arange(target_dims) // obs_every) + 1) * obs_every - 1, target_dims - 1)
    obs = noisy[:, nxt]                                   # every step carries the NEXT observation (look-ahead fill)
    time_till = np.tile((nxt - np.arange(target_dims)) * dt, (2, 1))
    return obs, obs_bin, time_till


def generate(target_dims=1000000, dt=0.1, x0=(2., 3.), obs_every=10, obs_std=0.1, seed=1, dat_dir="dat"):
    """Writes the three files fitz_nag_NVP.py:452-454 loads (the reference repository does not ship them)."""
    obs, obs_bin, time_till = simulate(target_dims, dt, x0, obs_every, obs_std, seed)
    os.makedirs(dat_dir, exist_ok=True)
    np.savetxt(os.path.join(dat_dir, "fitz_nag_obs_partial.txt"), obs)
    np.savetxt(os.path.join(dat_dir, "fitz_nag_obs_binary.txt"), obs_bin)
    np.savetxt(os.path.join(dat_dir, "fitz_nag_time_till.txt"), time_till)
    return obs, obs_bin, time_till


def main(train=False, p=50, kernel_len=20, dt=0.1, T=100000., batch_dims=50, network_dims=(50,) * 5, no_flows=3,
         priors=((0., 10.),) * 5, feat_window=10, x0=(2., 3.), learn_rate=1e-4, early_stopping=None):
    """fitz_nag_NVP.py:451-523."""
    obs = np.loadtxt('dat/fitz_nag_obs_partial.txt', NP_DTYPE)
    obs_bin = np.loadtxt('dat/fitz_nag_obs_binary.txt', NP_DTYPE)
    time_till = np.loadtxt('dat/fitz_nag_time_till.txt', NP_DTYPE)
    target_dims = int(np.int32(T / dt))
    if obs.shape[1] != target_dims:                        # a shorter generated series: keep T consistent with it
        target_dims = obs.shape[1]
        T = target_dims * dt
    theta_dist = ThetaFlow(len(priors), 4, base_loc=0., base_scale=1., activation="elu")
    if early_stopping is None:
        early_stopping = float(os.environ.get("NMA_MAX_STEPS", "1e99"))
    var_model = VI_SSM(obs, obs_bin, time_till, np.array(x0), theta_dist, list(priors), dt, T, p, kernel_len,
                       batch_dims, list(network_dims), target_dims, no_flows, feat_window, learn_rate=learn_rate,
                       pre_train=True, early_stopping=early_stopping)
    var_model.build_flow()
    os.makedirs('locally_variant', exist_ok=True)
    if train:
        var_model.train(tensorboard_path='locally_variant/train/',
                        save_path='model_saves/fitz_nag_model_%i.ckpt' % batch_dims)
    var_model.save_paths('locally_variant/FHN_obs_paths.txt')
    import torch
    with torch.no_grad():
        z0 = theta_dist.base_sample(100000, None, var_model.device)
        np.savetxt('locally_variant/FHN_local_post.txt', theta_dist.sample_and_log_prob(z0)[0].cpu().numpy())
    return var_model


if __name__ == "__main__":
    if "--generate" in sys.argv:
        k = sys.argv.index("--generate")
        generate(int(sys.argv[k + 1]) if len(sys.argv) > k + 1 else 1000000)
    else:
        main(train="--train" in sys.argv)
