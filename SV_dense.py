"""Stochastic-volatility NMA model - drop-in for the reference's SV_dense.py.

`from SV_dense import VI_SSM` gives the class with the reference's constructor and methods (SV_dense.py:139-402);
`python SV_dense.py` runs what the bottom of the reference script runs (:405-463): load dat/SV.dat (rows from 300
on), build the theta posterior (5 inverse-MAF layers, base N(0, 1), relu, 4 permutations from numpy's global
stream), build the model, write theta samples and posterior paths under locally_variant/, train.  The TensorFlow
graph is replaced by the B200 library (viforssms_b200/vi_ssm_models.py).

dat/SV.dat is the reference's own data file (one price per line); it is an input, not part of this repository.
`python SV_dense.py --generate [N]` writes a synthetic series from the script's own SDE (SV_dense.py:211-223 at
the pre-training target theta, x0 = -8.5) in the same one-column layout.
"""
import os
import sys

import numpy as np

from viforssms_b200.theta_flow import ThetaFlow
from viforssms_b200.vi_ssm_models import SV_VI_SSM as VI_SSM

NP_DTYPE = np.float32
np.random.seed(1)                      # SV_dense.py:18

__all__ = ["VI_SSM", "main", "generate", "simulate", "ThetaFlow", "NP_DTYPE"]

THETA_STAR = (0.001, -.6, np.log(0.08), np.log(0.5))      # SV_dense.py:254


def simulate(n=1809, dt=1.0, x0=-8.5, s0=100.0, seed=1, device=None):
    """Euler-Maruyama of the two-component SDE of SV_dense.py:211-223: price S with drift theta0*S and diffusion
    S*exp(V/2), log-volatility V with drift theta1 - exp(theta2) V and diffusion exp(theta3).  Both recursions are
    affine in their state (V: constant coefficient; S: multiplicative), so with `device` given they run as the library's
    affine-map prefix scans (nma_scan_affine, the A12 kernel) instead of the Python loop."""
    rs = np.random.RandomState(seed)
    if device is not None:
        import torch
        from viforssms_b200.engine import scan_affine
        th = THETA_STAR
        z = torch.from_numpy(rs.standard_normal((2, n))).to(device)
        ones = torch.ones(n - 1, dtype=torch.float64, device=device)
        v = scan_affine(ones * (1.0 - dt * np.exp(th[2])), dt * th[1] + np.sqrt(dt) * np.exp(th[3]) * z[1, :n - 1], x0)
        mult = 1.0 + dt * th[0] + np.sqrt(dt) * torch.exp(0.5 * v[:n - 1]) * z[0, :n - 1]
        return scan_affine(mult.contiguous(), torch.zeros_like(mult), s0).cpu().numpy()
    th = THETA_STAR
    s, v = np.empty(n), np.empty(n)
    s[0], v[0] = s0, x0
    z = rs.standard_normal((2, n))
    for t in range(n - 1):
        s[t + 1] = s[t] + dt * th[0] * s[t] + np.sqrt(dt) * s[t] * np.exp(0.5 * v[t]) * z[0, t]
        v[t + 1] = v[t] + dt * (th[1] - np.exp(th[2]) * v[t]) + np.sqrt(dt) * np.exp(th[3]) * z[1, t]
    return s


def generate(n=1809, dt=1.0, x0=-8.5, s0=100.0, seed=1, path="dat/SV.dat", device=None):
    """Writes dat/SV.dat (SV_dense.py:406), which the reference repository does not ship."""
    s = simulate(n, dt, x0, s0, seed, device)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    np.savetxt(path, s)
    return s


def main(p=200, kernel_len=50, dt=1.0, batch_dims=52, network_dims=(50,) * 5, no_flows=5,
         priors=((0., 10.0),) * 4, feat_window=5, x0=-8.5, learn_rate=1e-4, early_stopping=None, train=True):
    """SV_dense.py:405-463."""
    obs = np.loadtxt('dat/SV.dat', NP_DTYPE)[300:]
    T = obs.shape[0] - 1
    target_dims = int(np.int32(T / dt))
    theta_dist = ThetaFlow(len(priors), 5, base_loc=0., base_scale=1., activation="relu")
    if early_stopping is None:
        early_stopping = float(os.environ.get("NMA_MAX_STEPS", "1e99"))
    var_model = VI_SSM(obs, x0, theta_dist, list(priors), dt, T, p, kernel_len, batch_dims, list(network_dims),
                       target_dims, no_flows, feat_window, learn_rate=learn_rate, pre_train=True,
                       early_stopping=early_stopping)
    var_model.build_flow()
    os.makedirs('locally_variant', exist_ok=True)
    import torch
    with torch.no_grad():
        z0 = theta_dist.base_sample(100000, None, var_model.device)
        np.savetxt('locally_variant/SV_local_post.txt', theta_dist.sample_and_log_prob(z0)[0].cpu().numpy())
    var_model.save_paths('locally_variant/SV_obs_paths.txt')
    if train:
        var_model.train(tensorboard_path='locally_variant/train/',
                        save_path='model_saves/SV_model_%i_v211.ckpt' % batch_dims)
    return var_model


if __name__ == "__main__":
    if "--generate" in sys.argv:
        k = sys.argv.index("--generate")
        generate(int(sys.argv[k + 1]) if len(sys.argv) > k + 1 else 1809)
    else:
        main()
