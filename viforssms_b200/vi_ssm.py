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


def restore_theta_perms(model, ck) -> None:
    """The theta posterior's fixed permutations are drawn from numpy's global stream when the model is built
    (AR.py:384-385): a checkpoint is only meaningful with the permutations it was trained with, so re-bind them."""
    perms = ck.get("perms")
    if perms is None:
        return
    flow = model.theta_dist
    new = [np.asarray(pm, dtype=np.int64) for pm in perms]
    if len(new) != len(flow.perms) or any(len(a) != flow.d for a in new):
        raise ValueError("checkpoint was written by a theta posterior of a different shape")
    if any(not np.array_equal(a, b) for a, b in zip(new, flow.perms)):
        flow.perms = new
        if flow.flat is not None:
            flow.bind(flow.flat)
        if model.eng is not None:
            model.eng.set_theta_flow(flow, model.priors)


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
        self._scalars_host = {}
        self.prior_t = prior_tensors(self.priors, self.device)
        # the whole iteration is one nma_train_step call (in-library noise, theta posterior, ELBO + gradients, clip +
        # Adamax, logged means); NMA_HOST_THETA=1 keeps the host autograd theta posterior (comparison path)
        self.eng.set_theta_flow(self.theta_dist, self.priors)
        self.eng.set_seed(self.seed, 0)
        self.scalars_dev = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._theta_last = torch.zeros(self.p, cfg.dtheta, dtype=torch.float32, device=self.device)
        self._host_theta = os.environ.get("NMA_HOST_THETA") == "1"
        self._use_graph = os.environ.get("NMA_FACADE_GRAPH", "1") != "0" and not self._host_theta
        self._graphs, self._seen = {}, set()

    # ------------------------------------------------------------------
    def _iteration(self, batch_select: np.ndarray, pre_train: bool) -> None:
        """One sess.run.  The body is a fixed sequence of stream-ordered C-ABI calls, so after one eager iteration per
        optimiser it is captured into a CUDA graph and replayed: at p = 50 the iteration is launch-bound otherwise.
        Every call trains (the capturing call replays the graph it has just recorded)."""
        self.idx_dev.copy_(torch.from_numpy(np.ascontiguousarray(batch_select, dtype=np.int64)), non_blocking=True)
        self._run(pre_train)

    def _run(self, pre_train: bool) -> None:
        """The iteration on the subsequence starts already in `idx_dev`."""
        if not self._use_graph:
            self._body(pre_train)
            return
        g = self._graphs.get(pre_train)
        if g is None:
            if pre_train not in self._seen:           # first call of this optimiser: eager
                self._seen.add(pre_train)
                self._body(pre_train)
                return
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._body(pre_train)
            self._graphs[pre_train] = g
        g.replay()

    def _body(self, pre_train: bool) -> None:
        """The body of one sess.run: sample theta and eps, ELBO + gradients, clip, Adamax, summaries."""
        if self._host_theta:
            return self._body_host_theta(pre_train)
        if pre_train:       # AdamaxOptimizer(1e-3, beta1=0.9).minimize(-obs_loss): no clipping (AR.py:201-202)
            m, v = self.slots["pre"]
            self.eng.train_step(self.blob, self.grad, m, v, self.idx_dev, self.scalars_dev, objective=OBJ_NEG_OBS,
                                prior_on=False, lr=1e-3, beta1=0.9, clip=0.0, theta_out=self._theta_last)
        else:               # clip_by_global_norm + AdamaxOptimizer(lr, beta1=0.95) (AR.py:226-234)
            m, v = self.slots["main"]
            self.eng.train_step(self.blob, self.grad, m, v, self.idx_dev, self.scalars_dev, objective=OBJ_ELBO,
                                prior_on=True, lr=self.learn_rate, beta1=0.95, clip=self.grad_clip,
                                theta_out=self._theta_last)

    TAGS = ("loss/ELBO", "loss/SDE_log_prob", "loss/theta_log_prob", "loss/obs_log_prob", "loss/path_log_prob",
            "optimize/global_norm")

    @property
    def scalars(self) -> dict:
        return self.read_scalars()

    def read_scalars(self) -> dict:
        """The summaries of the last iteration (AR.py:207-224) - ONE device-to-host copy."""
        if self._host_theta:
            return {k: float(v) for k, v in self._scalars_host.items()}
        vals = self.scalars_dev.cpu().tolist()
        return dict(zip(self.TAGS, vals[:6]))

    def _body_host_theta(self, pre_train: bool) -> None:
        """The same iteration with the host autograd theta posterior and torch.randn noise (comparison path)."""
        cfg = self.cfg
        gen = self.gen
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
        if pre_train:
            m, v = self.slots["pre"]
            self.eng.adamax_step(self.blob, self.grad, m, v, 1e-3, 0.9, clip=0.0)
        else:
            m, v = self.slots["main"]
            norm = self.eng.adamax_step(self.blob, self.grad, m, v, self.learn_rate, 0.95, clip=self.grad_clip)
            t = out["terms"]
            scale = float(cfg.scale)
            elbo = scale * (t[:, 0] - t[:, 2] + t[:, 1]) + prior.detach() - logq_theta.detach()
            self._scalars_host = {
                "loss/ELBO": elbo.mean(), "loss/SDE_log_prob": scale * t[:, 0].mean(),
                "loss/theta_log_prob": logq_theta.detach().mean(), "loss/obs_log_prob": scale * t[:, 1].mean(),
                "loss/path_log_prob": scale * t[:, 2].mean(), "optimize/global_norm": norm.clone()[0],
            }
            self._theta_last = theta.detach()
        if self.theta_dist.tf_mask_grad:
            self.theta_dist.constrain()

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
                    for tag, val in self.read_scalars().items():
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
        restore_theta_perms(self, ck)
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
