"""AR(1) series generator with the reference's interface and byte-identical output files.

Drop-in for the reference's `AR_dat_gen.py:6-43`: `data_gen(T, impute, x0, theta, obs_std, dat_dir)`
writes dat/AR_obs_partial.txt, dat/AR_obs_binary.txt and dat/AR_time_till.txt with `np.savetxt`.
The module seeds numpy's legacy global stream with 1 at import, like the reference (AR_dat_gen.py:3),
and consumes it in the same order (T scalar normals for the latent path, then T+1 for the
observations), so `data_gen(5000, 1, 10.0, [5, .5, 3], 1.)` reproduces the committed dat/AR_*.txt
byte for byte (tests/test_oracle.py checks the sha256 recorded from the reference's files).

For long series (T >= 10^6) `data_gen_device` runs the recurrence as an affine prefix scan on the GPU
(viforssms_b200.engine.scan_ar1 / time_till); it draws the same normals but reassociates the
recurrence, so it agrees with the loop to ~1e-12 relative, not bit for bit.
"""
import os

import numpy as np

np.random.seed(1)


def simulate(T, impute, x0, theta, obs_std):
    """Returns (obs_fill, obs_binary, time_till_out) as float64 arrays."""
    theta = np.asarray(theta, dtype=np.float64)
    n = int(np.int32(T + 1))
    # legacy_gauss stream: drawing n-1 standard normals at once consumes the stream exactly like n-1
    # scalar np.random.normal(loc, scale) calls, and normal(loc, scale) == loc + scale * gauss.
    z = np.random.standard_normal(n - 1)
    X = np.empty(n)
    X[0] = x0
    a, b, c = float(theta[1]), float(theta[0]), float(theta[2])
    prev = X[0]
    for i in range(1, n):
        prev = (prev * a + b) + c * z[i - 1]
        X[i] = prev
    obs = X + obs_std * np.random.standard_normal(n)

    kept = obs[impute:][0::impute]
    m = kept.shape[0] * impute
    obs_partial = np.zeros(m)
    obs_partial[impute - 1::impute] = kept
    obs_fill = np.repeat(kept, impute)
    obs_binary = (obs_partial != 0).astype(np.float64)
    # count-down to the next observation: distance from the last observed slot, reset at observations
    pos = np.arange(m)
    last = np.maximum.accumulate(np.where(obs_binary == 1.0, pos, -1))
    time_till = np.where(obs_binary == 1.0, 0.0, (pos - last).astype(np.float64))
    return obs_fill, obs_binary, -(time_till - impute)


def _write(dat_dir, obs_fill, obs_binary, time_till_out):
    d = os.path.join(dat_dir, "dat")
    if not os.path.exists(d):
        os.makedirs(d)
    for name, arr in (("AR_obs_partial.txt", obs_fill), ("AR_obs_binary.txt", obs_binary),
                      ("AR_time_till.txt", time_till_out)):
        with open(os.path.join(d, name), "w+") as f:
            np.savetxt(f, arr)


def data_gen(T, impute, x0, theta, obs_std, dat_dir=os.getcwd()):
    theta = np.asarray(theta, dtype=np.float64)   # main.py -t hands over strings (main.py:116-118)
    _write(dat_dir, *simulate(T, impute, x0, theta, obs_std))


def data_gen_device(T, impute, x0, theta, obs_std, seed=1, device="cuda"):
    """Same series on the GPU (float64 tensors): A12/A13 of SURVEY §8a.  Returns device tensors."""
    import torch
    from viforssms_b200.engine import scan_ar1, time_till
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = int(T)
    z = torch.randn(n, dtype=torch.float64, device=device, generator=g)
    X = scan_ar1(z, float(x0), float(theta[1]), float(theta[0]), float(theta[2]))
    obs = X + float(obs_std) * torch.randn(n + 1, dtype=torch.float64, device=device, generator=g)
    return time_till(obs.contiguous(), int(impute))


if __name__ == "__main__":
    data_gen(T=5000, impute=1, x0=10.0, theta=np.array([5.0, .5, 3.0]), obs_std=1.)
