"""Command-line entry point with the reference's surface (main.py:6-130):

    python main.py hyperparameters.txt [-T N] [-i N] [-t v ...] [-x v] [-o v] [-k N] [-b N] [-f N] [-repair]

The hyper-parameter file is read by fixed line position (values on lines 2, 4, ..., 28); command-line options
override it; the series is (re)generated into dat/ and the model trains on it.
"""
import argparse
import sys

import numpy as np

from AR_dat_gen import data_gen
from AR import main

# (header, default) pairs in file order; values sit on the even lines (1-based) of the hyper-parameter file
FIELDS = [
    ("T  (number of time steps of the series)", "5000"),
    ("impute  (keep every impute-th observation)", "1"),
    ("x0  (latent value at time 0)", "10.0"),
    ("theta  (theta0, theta1, theta2 of x_t = theta1 x_{t-1} + theta0 + theta2 xi_t)", "5.0, 0.5, 3.0"),
    ("obs_std  (observation noise standard deviation)", "1."),
    ("p  (MC samples = subsequences per iteration)", "50"),
    ("kernel_len  (taps of the moving-average conv)", "50"),
    ("batch_dims  (latent steps per subsequence)", "50"),
    ("network_dims", "50, 50, 50"),
    ("no_flows", "3"),
    ("priors  ((mean, scale) per theta)", "(0., 10.0)(0., 10.0)(0., 10.0)"),
    ("feat_window  (look-ahead observations fed as features)", "10"),
    ("learn_rate", "1e-3"),
    ("grad_clip  (global-norm clip)", "2.5e8"),
]
DEFAULT_FILE = "\n".join("#### %s ####\n%s" % (h, v) for h, v in FIELDS) + "\n"

# (short flag, long flag, destination, how it accumulates)
OPTIONS = [("-T", "-time", "T", "store"), ("-i", "-impute", "impute", "store"), ("-t", "-theta", "theta", "append"),
           ("-x", "-xzero", "x0", "store"), ("-o", "-obs_std", "obs_std", "store"),
           ("-k", "-kernel_len", "kernel_len", "store"), ("-b", "-batch_dims", "batch_dims", "store"),
           ("-f", "-feat_window", "feat_window", "store")]


def handle_opts(argv=None):
    parser = argparse.ArgumentParser(
        usage="%(prog)s hyperparameters.txt [OPTIONS]\n  options given on the command line win over the file; "
              "-repair prints a default file to copy into hyperparameters.txt")
    parser.add_argument("file", help="hyper-parameter file (values on every second line, fixed order)")
    for short, long_, dest, action in OPTIONS:
        parser.add_argument(short, long_, action=action, dest=dest, default=None,
                            help="override %s (repeat -t once per theta component)" % dest)
    parser.add_argument("-repair", action="store_true", dest="repair", default=False,
                        help="print the default hyper-parameter file and exit")
    return parser.parse_args(argv)


def parseparams(file):
    """Values sit on the odd lines (0-based 1, 3, ..., 27) in a fixed order (main.py:26-57)."""
    with open(file, "r") as f:
        v = [ln.rstrip() for ln in f.readlines()]
    pairs = v[21].replace(')', '').split("(")[1:]
    return [int(v[1]), int(v[3]), float(v[5]), [float(t) for t in v[7].split(",")], float(v[9]), int(v[11]),
            int(v[13]), int(v[15]), [int(d) for d in v[17].split(",")], int(v[19]),
            [(float(t.split(",")[0]), float(t.split(",")[1])) for t in pairs], int(v[23]), float(v[25]), float(v[27])]


def resolve(args):
    try:
        (T, impute, x0, theta, obs_std, p, kernel_len, batch_dims, network_dims, no_flows, priors, feat_window,
         learn_rate, grad_clip) = parseparams(args.file)
        theta = np.array(theta)
    except Exception:
        sys.exit("Please specify a valid hyperparameter file")
    if args.T is not None:
        T = int(args.T)
    if args.impute is not None:
        impute = int(args.impute)
    if args.theta is not None:
        theta = np.array(args.theta)          # strings, like the reference (main.py:116-118); data_gen converts
    if args.x0 is not None:
        x0 = float(args.x0)
    if args.obs_std is not None:
        obs_std = float(args.obs_std)
    if args.kernel_len is not None:
        kernel_len = int(args.kernel_len)
    if args.batch_dims is not None:
        batch_dims = int(args.batch_dims)
    if args.feat_window is not None:
        feat_window = int(args.feat_window)
    return dict(T=T, impute=impute, x0=x0, theta=theta, obs_std=obs_std, p=p, kernel_len=kernel_len,
                batch_dims=batch_dims, network_dims=network_dims, no_flows=no_flows, priors=priors,
                feat_window=feat_window, learn_rate=learn_rate, grad_clip=grad_clip)


if __name__ == "__main__":
    args = handle_opts()
    if args.repair:
        print(DEFAULT_FILE)
        sys.exit("Copy the above into a .txt file")
    h = resolve(args)
    data_gen(h["T"], h["impute"], h["x0"], h["theta"], h["obs_std"])
    main(h["p"], h["kernel_len"], h["T"], h["batch_dims"], h["network_dims"], h["no_flows"], h["priors"],
         h["feat_window"], h["x0"], h["obs_std"], learn_rate=h["learn_rate"], grad_clip=h["grad_clip"])
