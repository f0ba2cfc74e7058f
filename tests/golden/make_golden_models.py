#!/usr/bin/env python
"""Golden feeds of the FitzHugh-Nagumo, stochastic-volatility and Lotka-Volterra (fixed theta) scripts, produced by
the reference's own unmodified code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_models.py

fitz_nag_NVP.py, SV_dense.py and lotka_volterra_partial_batch_fix_theta.py are module-level scripts: importing them
loads the series, builds the TensorFlow graph, opens a session and calls save_paths()/train().  TensorFlow 1.8
cannot run here, but every line that decides WHAT is fed to the graph is numpy: series padding
(fitz_nag_NVP.py:187-202, SV_dense.py:159-184, lotka_volterra_partial_batch_fix_theta.py:203-222), subsequence
sampling (np.random.choice) and the window gather.  We exec each script's source under a stub `tensorflow` whose
session records every feed_dict, with np.loadtxt patched to hand the script the synthetic series of
tests/golden/synth.py, and store what the script fed: the first two save_paths() feeds and the first two train()
feeds.  Arrays too large to commit are stored as their first rows plus a sha256 of the whole array.

Outputs: fhn_golden.npz, sv_golden.npz, lv_golden.npz, lvb_golden.npz
"""
import hashlib
import os
import sys
import tempfile
from unittest import mock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402


class _Stop(BaseException):     # not an Exception: the LV script wraps its body in `except Exception`
    pass


class _Session:
    def __init__(self):
        self.feeds = []
        self.limit = 10 ** 9
        self.graph = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        if feed_dict is not None:
            self.feeds.append(dict(feed_dict))
            if len(self.feeds) >= self.limit:
                raise _Stop()
        if isinstance(fetches, (list, tuple)):
            return [np.zeros(1) for _ in fetches]
        return np.zeros((1, 1, 2))


def _install_tf_stub(session):
    tf = mock.MagicMock(name="tensorflow")
    tf.float32 = "float32"
    tf.InteractiveSession = lambda *a, **k: session
    tf.Session = lambda *a, **k: session
    counter = {"n": 0}

    def placeholder(*a, **k):
        counter["n"] += 1
        return mock.MagicMock(name="placeholder%d" % counter["n"])
    tf.placeholder = placeholder

    class _Base(object):        # base of the scripts' own subclasses (init_dist(tfd.Normal), AdamaxOptimizer)
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return mock.MagicMock(name=name)

        def compute_gradients(self, *a, **k):
            return [(mock.MagicMock(), mock.MagicMock())]
    tf.split = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
    tf.clip_by_global_norm = lambda g, c: (list(g), mock.MagicMock())
    tf.contrib.distributions.Normal = _Base
    tf.python.training.optimizer.Optimizer = _Base
    for name in ("tensorflow", "tensorflow.python", "tensorflow.python.ops", "tensorflow.python.ops.clip_ops",
                 "tensorflow.python.framework", "tensorflow.python.framework.ops", "tensorflow.python.training",
                 "tensorflow.python.training.optimizer", "tensorflow.contrib", "tensorflow.contrib.distributions",
                 "tensorflow.python.client"):
        obj = tf
        for part in name.split(".")[1:]:
            obj = getattr(obj, part)
        sys.modules[name] = obj
    # plotting is imported by the scripts but never reached on the paths we execute
    plt = mock.MagicMock(name="matplotlib")
    sys.modules.setdefault("matplotlib", plt)
    sys.modules.setdefault("matplotlib.pyplot", plt.pyplot)
    return tf


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_script(script, arrays, dirs, n_paths_feeds, after):
    """exec `script` (a file of the reference) with np.loadtxt returning `arrays` in order; stop after
    `n_paths_feeds` recorded feeds; then `after(ns, sess)` drives train().  Returns (namespace, feeds, draws)."""
    sess = _Session()
    _install_tf_stub(sess)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    work = tempfile.mkdtemp()
    os.chdir(work)
    for d in dirs:
        os.makedirs(os.path.join(work, d), exist_ok=True)
    queue = list(arrays)
    real_open = open

    def fake_loadtxt(*a, **k):
        return np.array(queue.pop(0))

    def fake_open(path, mode="r", *a, **k):
        if isinstance(path, str) and path.startswith("dat/") and "r" in mode:
            return real_open(os.devnull, "r")
        return real_open(path, mode, *a, **k)
    draws = []
    real_choice = np.random.choice

    def rec_choice(*a, **k):
        out = real_choice(*a, **k)
        draws.append(np.array(out))
        return out
    with real_open(os.path.join(REF, script)) as f:
        src = f.read()
    ns = {"__name__": "reference_script", "__file__": os.path.join(REF, script)}
    sess.limit = n_paths_feeds
    with mock.patch("numpy.loadtxt", fake_loadtxt), mock.patch("numpy.savetxt"), mock.patch("numpy.save"), \
            mock.patch("builtins.open", fake_open), mock.patch("numpy.random.choice", rec_choice):
        try:
            exec(compile(src, os.path.join(REF, script), "exec"), ns)
        except _Stop:
            pass
        n0 = len(sess.feeds)
        try:
            after(ns, sess)
        except _Stop:
            pass
    return ns, sess.feeds, draws, n0


def pack(out, tag, feed, model, extra_names):
    tf_feed = np.asarray(feed[model.time_feats])
    assert tf_feed.dtype == np.float64
    out[tag + "_time_feats_shape"] = np.array(tf_feed.shape)
    out[tag + "_time_feats_sha256"] = np.array(sha(tf_feed))
    out[tag + "_time_feats_f32_sha256"] = np.array(sha(tf_feed.astype(np.float32)))
    out[tag + "_time_feats_rows"] = tf_feed[:3]
    out[tag + "_mask"] = np.asarray(feed[model.mask])
    out[tag + "_shift"] = np.asarray(feed[model.shift])
    for name in extra_names:
        out[tag + "_" + name] = np.asarray(feed[getattr(model, name)])


def main():
    # ---------------- FitzHugh-Nagumo ----------------
    obs, obs_bin, tt = synth.fhn_inputs()

    def fhn_after(ns, sess):
        np.random.seed(101)
        sess.limit = len(sess.feeds) + 2
        ns["var_model"].train(tensorboard_path="locally_variant/train/", save_path="model_saves/x.ckpt")
    ns, feeds, draws, n0 = run_script("fitz_nag_NVP.py", [obs, obs_bin, tt], ["dat", "locally_variant", "model_saves"],
                                      2, fhn_after)
    m = ns["var_model"]
    out = {"hyper": np.array([m.p, m.kernel_len, m.batch_dims, m.no_flows, int(m.target_dims), 10]),
           "dt": np.array(m.dt), "T": np.array(m.T), "x0": np.array(ns["x0"])}
    assert n0 == 2 and len(feeds) == 4 and len(draws) == 2
    for k in range(2):
        pack(out, "paths%d" % k, feeds[k], m, ["bin_feed"])
        out["paths%d_batch_select" % k] = np.tile(k * m.batch_dims, m.p)
        pack(out, "train%d" % k, feeds[2 + k], m, ["bin_feed"])
        out["train%d_batch_select" % k] = draws[k]
    np.savez_compressed(os.path.join(HERE, "fhn_golden.npz"), **out)
    print("fhn:", {k: getattr(v, "shape", None) for k, v in out.items()})

    # ---------------- stochastic volatility ----------------
    prices = synth.sv_prices()

    def sv_after(ns, sess):
        np.random.seed(202)
        sess.limit = len(sess.feeds) + 2
        ns["var_model"].train(tensorboard_path="locally_variant/train/", save_path="model_saves/x.ckpt")
    ns, feeds, draws, n0 = run_script("SV_dense.py", [prices], ["dat", "locally_variant", "model_saves"], 2, sv_after)
    m = ns["var_model"]
    out = {"hyper": np.array([m.p, m.kernel_len, m.batch_dims, m.no_flows, int(m.target_dims), 5]),
           "dt": np.array(m.dt), "T": np.array(ns["T"]), "x0": np.array(ns["x0"]),
           "var_pad": m.var_pad, "var_diff_pad": m.var_diff_pad}
    assert n0 == 2 and len(feeds) == 4 and len(draws) == 2
    for k in range(2):
        pack(out, "paths%d" % k, feeds[k], m, ["dim_one"])
        out["paths%d_batch_select" % k] = np.tile(k * m.batch_dims, m.p)
        pack(out, "train%d" % k, feeds[2 + k], m, ["dim_one"])
        out["train%d_batch_select" % k] = draws[k]
    np.savez_compressed(os.path.join(HERE, "sv_golden.npz"), **out)
    print("sv:", {k: getattr(v, "shape", None) for k, v in out.items()})

    # ---------------- Lotka-Volterra, fixed theta (series 0 of the concatenated file) ----------------
    obs, obs_bin, tt = synth.lv_inputs()
    ns, feeds, draws, n0 = run_script(
        "lotka_volterra_partial_batch_fix_theta.py", [obs, obs_bin, tt],
        ["dat/our_files/fix_theta", "locally_variant/fix_theta/train_dense", "model_saves/fix_theta"], 3,
        lambda ns, sess: None)
    m = ns["var_model"]
    out = {"hyper": np.array([m.p_val, m.kernel_len, m.batch_dims, m.no_flows, int(m.target_dims), 10]),
           "dt": np.array(m.dt), "T": np.array(ns["T"]), "x0_mean": np.array(ns["x0_mean"]),
           "priors": np.array(ns["priors"]), "obs_not_observed": np.array(ns["obs_not_observed"]),
           "time_feats_full": np.asarray(feeds[0][m.time_feats])}
    assert len(feeds) == 3 and len(draws) == 2
    pack(out, "paths0", feeds[0], m, ["bin_feed"])
    out["paths0_batch_select"] = np.tile(0, m.p_val)
    for k in range(2):
        pack(out, "train%d" % k, feeds[1 + k], m, ["bin_feed"])
        out["train%d_batch_select" % k] = draws[k]
    np.savez_compressed(os.path.join(HERE, "lv_golden.npz"), **out)
    print("lv:", {k: getattr(v, "shape", None) for k, v in out.items()})
    lv_batch()


def lv_batch():
    """lotka_volterra_partial_batch.py as committed: p_val = 3 windows per iteration over the first three series of the
    concatenated file (:677-764); 3 save_paths feeds, then train feeds."""
    obs, obs_bin, tt = synth.lv_inputs()
    ns, feeds, draws, n0 = run_script("lotka_volterra_partial_batch.py", [obs, obs_bin, tt],
                                      ["dat/our_files", "locally_variant/train", "model_saves"], 6, lambda ns, sess: None)
    m = ns["var_model"]
    out = {"hyper": np.array([m.p_val, m.kernel_len, m.batch_dims, m.no_flows, int(m.target_dims), 10]),
           "dt": np.array(m.dt), "T": np.array(ns["T"]), "x0_mean": np.array(ns["x0_mean"]),
           "obs_not_observed": np.array(ns["obs_not_observed"])}
    assert len(feeds) == 6 and len(draws) == 3 and m.p_val == 3
    for k in range(3):
        pack(out, "paths%d" % k, feeds[k], m, ["bin_feed"])
        pack(out, "train%d" % k, feeds[3 + k], m, ["bin_feed"])
        out["train%d_batch_select" % k] = draws[k]
    np.savez_compressed(os.path.join(HERE, "lvb_golden.npz"), **out)
    print("lv batch:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    if "--only-lvb" in sys.argv:
        lv_batch()
    else:
        main()
