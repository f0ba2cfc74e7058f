"""`VI_SSM` — the reference's model/trainer class for the AR(1) NMA model (AR.py:113-362), re-hosted on
the B200 library.

Same constructor arguments, same methods (`build_flow`, `train`, `save`, `load`, `save_paths`), same
side effects (TensorBoard scalars under the same tags, a checkpoint every 1000 iterations, posterior
paths written with `np.savetxt`).  What the reference does inside `sess.run([train_step, merged],
feed_dict)` (AR.py:300-301) happens here in `nma_elbo_fwd_bwd` + `nma_adamax_step`; the per-iteration
numpy window gather (AR.py:267-288) is replaced by the device gather, fed only the subsequence starts,
which are still drawn with the reference's own `np.random.choice` call on numpy's global legacy stream.
"""
from __future__ import annotations

import os
import time
from datetime import datetime
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import feed
from .config import OBJ_ELBO, OBJ_NEG_OBS, ar_config, param_layout
from .engine import NMAEngine
from .theta_flow import ThetaFlow, prior_log_prob, prior_tensors
from .trainer import glorot_blob


class VI_SSM:
    def __init__(self, obs, obs_std, x0, theta_dist: ThetaFlow, priors: Sequence[Tuple[float, float]], T, p,
                 kernel_len, batch_dims, network_dims, no_flows, feat_window, obs_bin, time_till, pre_train=False,
                 early_stopping=1e99, learn_rate=1e-3, grad_clip=2.5e8, device: Optional[torch.device] = None,
                 seed: int = 1):
        if len(set(network_dims)) != 1:
            raise ValueError("all network_dims must be equal (the reference only ever uses [50]*n)")
        self.priors = list(priors)
        self.obs_std = obs_std
        self.T = np.int32(T)
        self.p = int(p)
        self.kernel_len = int(kernel_len)
        self.batch_dims = int(batch_dims)
        self.network_dims = list(network_dims)
        self.no_flows = int(no_flows)
        self.feat_window = int(feat_window)
        self.pre_train = pre_train
        self.early_stopping = early_stopping
        self.learn_rate = learn_rate
        self.grad_clip = grad_clip
        self.kernel_ext = self.kernel_len * self.no_flows + self.batch_dims + 1     # AR.py:132
        self.theta_dist = theta_dist
        self.x0 = float(x0)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.seed = seed
        self._series = (np.asarray(obs), np.asarray(obs_bin), np.asarray(time_till))
        self.cfg = ar_config(p=self.p, K=self.kernel_len, B=self.batch_dims, F=self.no_flows,
                             H=len(self.network_dims) - 2, feat_window=self.feat_window, T=int(self.T),
                             obs_std=float(obs_std), x0=self.x0)
        self.cfg.C = int(self.network_dims[0])
        self.eng: Optional[NMAEngine] = None

    # ------------------------------------------------------------------
    def build_flow(self) -> None:
        """Model assembly (AR.py:189-238): engine + series on the device, variables, the two optimisers' slots."""
        cfg = self.cfg
        self.eng = NMAEngine(cfg, self.device)
        obs, obs_bin, tt = self._series
        self.eng.set_series(feed.ar_base_arrays(obs, obs_bin, tt, int(self.T), cfg.F, cfg.K, self.feat_window))
        g = torch.Generator().manual_seed(self.seed)
        _, self.n_nma = param_layout(cfg)
        nma = glorot_blob(cfg, g)
        self.blob = torch.cat([nma, self.theta_dist.init_values(g)]).to(self.device)
        self.n_total = self.blob.numel()
        # separate Adamax slots for the pre-train and the main optimiser (AR.py:201,227)
        self.slots = {"pre": (torch.zeros_like(self.blob), torch.zeros_like(self.blob)),
                      "main": (torch.zeros_like(self.blob), torch.zeros_like(self.blob))}
        self.grad = torch.zeros_like(self.blob)
        self.theta_leaf = self.blob[self.n_nma:].detach().requires_grad_(True)
        self.theta_dist.bind(self.theta_leaf)
        self.out = self.eng.alloc_outputs(self.p)
        self.out["grad_params"] = self.grad[:self.n_nma]
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(self.seed)
        self.idx_dev = torch.empty(self.p, dtype=torch.int64, device=self.device)
        self.scalars = {}
        self.prior_t = prior_tensors(self.priors, self.device)

    # ------------------------------------------------------------------
    def _iteration(self, batch_select: np.ndarray, pre_train: bool) -> None:
        """One sess.run.  With NMA_FACADE_GRAPH=1 the body is captured once per optimiser into a CUDA graph and
        replayed (at p = 50 the iteration is ~460 tiny launches, mostly the theta posterior: launch-bound); the
        capture path was written without a GPU at hand and is therefore opt-in."""
        self.idx_dev.copy_(torch.from_numpy(np.ascontiguousarray(batch_select, dtype=np.int64)))
        if os.environ.get("NMA_FACADE_GRAPH") != "1":
            self._body(pre_train, self.gen)
            return
        if not hasattr(self, "_graphs"):
            self._graphs, self._warm = {}, {True: 0, False: 0}
            torch.cuda.manual_seed(self.seed)              # captured randn draws come from the default CUDA generator
        g = self._graphs.get(pre_train)
        if g is None:
            if self._warm[pre_train] < 3:                  # eager warm-up on a side stream, as capture requires
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._body(pre_train, None)
                torch.cuda.current_stream().wait_stream(side)
                self._warm[pre_train] += 1
                return
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body(pre_train, None)
            self._graphs[pre_train] = g
            return
        g.replay()

    def _body(self, pre_train: bool, gen) -> None:
        """The body of one sess.run: sample theta and eps, ELBO + gradients, clip, Adamax."""
        cfg = self.cfg
        z0 = self.theta_dist.base_sample(self.p, gen, self.device)
        theta, logq_theta = self.theta_dist.sample_and_log_prob(z0)
        eps = torch.randn(self.p, cfg.L0, device=self.device, generator=gen)
        obj = OBJ_NEG_OBS if pre_train else OBJ_ELBO
        out = self.eng.elbo_fwd_bwd(self.blob[:self.n_nma], eps, theta.detach().contiguous(), self.idx_dev,
                                    objective=obj, out=self.out)
        prior = prior_log_prob(theta, self.prior_t)
        host_loss = (out["grad_theta"] * theta).sum()
        if not pre_train:
            host_loss = host_loss - (prior - logq_theta).sum()          # AR.py:184-185
        self.theta_leaf.grad = None
        host_loss.backward()
        self.grad[self.n_nma:].copy_(self.theta_leaf.grad)
        if pre_train:       # AdamaxOptimizer(1e-3, beta1=0.9).minimize(-obs_loss): no clipping (AR.py:201-202)
            m, v = self.slots["pre"]
            self.eng.adamax_step(self.blob, self.grad, m, v, 1e-3, 0.9, clip=0.0)
        else:               # clip_by_global_norm + AdamaxOptimizer(lr, beta1=0.95) (AR.py:226-234)
            m, v = self.slots["main"]
            norm = self.eng.adamax_step(self.blob, self.grad, m, v, self.learn_rate, 0.95, clip=self.grad_clip)
            t = out["terms"]
            scale = float(cfg.scale)
            elbo = scale * (t[:, 0] - t[:, 2] + t[:, 1]) + prior.detach() - logq_theta.detach()
            self.scalars = {
                "loss/ELBO": elbo.mean(), "loss/SDE_log_prob": scale * t[:, 0].mean(),
                "loss/theta_log_prob": logq_theta.detach().mean(), "loss/obs_log_prob": scale * t[:, 1].mean(),
                "loss/path_log_prob": scale * t[:, 2].mean(), "optimize/global_norm": norm.clone()[0],
            }
            self._theta_last = theta.detach()

    def _draw(self, replace_bool: bool) -> np.ndarray:
        sample_index = np.arange(0, self.T, self.batch_dims)
        return np.random.choice(sample_index, size=self.p, replace=replace_bool)     # AR.py:263-265

    def train(self, tensorboard_path, save_path, log_every: int = 1):
        if not os.path.exists(tensorboard_path):
            os.makedirs(tensorboard_path)
        save_parent_dir = os.path.dirname(save_path)
        if save_parent_dir and not os.path.exists(save_parent_dir):
            os.makedirs(save_parent_dir)
        writer = None
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter('%s/%s' % (tensorboard_path, datetime.now().strftime("%d:%m:%y-%H:%M:%S")))
        except Exception:       # tensorboard not installed: train without event files
            writer = None
        run = 0
        print("Training model...")
        converged = False
        replace_bool = bool(self.batch_dims * self.p >= self.T)
        theta_pos_index = [False, False, True]
        t_start = time.time()
        while not converged:
            batch_select = self._draw(replace_bool)
            if self.pre_train:
                if run == 0:
                    print("Pre-training...")
                self._iteration(batch_select, pre_train=True)
                if run == 500:
                    self.pre_train = False
                    print("Finished pre-training")
                    run = 0
            else:
                self._iteration(batch_select, pre_train=False)
                if writer is not None and run % log_every == 0:
                    for tag, val in self.scalars.items():
                        writer.add_scalar(tag, float(val), run)
                    th = self._theta_last
                    for i, pos in enumerate(theta_pos_index):
                        writer.add_histogram("parameters/%d" % i, (th[:, i].exp() if pos else th[:, i]).cpu(), run)
            if run == self.early_stopping:
                converged = True
            if run % 1000 == 0:
                self.save(save_path)
            run += 1
        if writer is not None:
            writer.close()
        self.train_seconds = time.time() - t_start

    # ------------------------------------------------------------------
    def save(self, PATH):
        torch.save({"blob": self.blob.cpu(), "slots": {k: (m.cpu(), v.cpu()) for k, (m, v) in self.slots.items()},
                    "numpy_rng": np.random.get_state(), "perms": [p.tolist() for p in self.theta_dist.perms]}, PATH)
        print("Model saved")

    def load(self, PATH):
        self.pre_train = False
        ck = torch.load(PATH, weights_only=False)
        self.blob.copy_(ck["blob"].to(self.device))
        for k, (m, v) in ck["slots"].items():
            self.slots[k][0].copy_(m.to(self.device))
            self.slots[k][1].copy_(v.to(self.device))
        print("Model restored")

    # ------------------------------------------------------------------
    def sample_paths(self, temp_index: int) -> torch.Tensor:
        """p posterior samples of the batch_dims+1 latent steps starting at `temp_index` (forward only)."""
        cfg = self.cfg
        idx = torch.full((self.p,), int(temp_index), dtype=torch.int64, device=self.device)
        with torch.no_grad():
            z0 = self.theta_dist.base_sample(self.p, self.gen, self.device)
            theta, _ = self.theta_dist.sample_and_log_prob(z0)
        eps = torch.randn(self.p, cfg.L0, device=self.device, generator=self.gen)
        _, lf = self.eng.forward_paths(self.blob[:self.n_nma], eps, theta.contiguous(), idx)
        return lf

    def save_paths(self, PATH_obs):
        """AR.py:323-362 (which refers to a placeholder the AR script never defines; the intent — every row
        evaluates the same subsequence, the windows are concatenated along time — is what fitz_nag_NVP.py:409-448
        does and what this does)."""
        path_stack = []
        for temp_index in np.arange(0, self.T, self.batch_dims):
            path_stack.append(self.sample_paths(int(temp_index))[:, 1:].cpu().numpy())
        paths = np.concatenate(path_stack, 1)
        with open(PATH_obs, 'w') as f:
            np.savetxt(f, paths)
        return paths
