"""`VI_SSM` facades of the FitzHugh-Nagumo (fitz_nag_NVP.py:158-448) and stochastic-volatility (SV_dense.py:139-402)
scripts, re-hosted on the B200 library.

Each class keeps the constructor arguments and the methods of the reference class of the same script
(`build_flow`, `train`, `save`, `load`, `save_paths`) and the same side effects (TensorBoard scalars under the same
tags, a checkpoint every 1000 iterations, posterior paths written with `np.savetxt` in the script's layout).  What
`sess.run([train_step, merged], feed_dict)` does in the reference (fitz_nag_NVP.py:389-390, SV_dense.py:341-342)
happens in `nma_elbo_fwd_bwd` + `nma_adamax_step`; the numpy window gather of every iteration is replaced by the
device gather, fed only the subsequence starts, which are still drawn with the reference's own `np.random.choice`
call on numpy's global legacy stream.

Pre-training runs the scripts' TWO optimisers in the same step (fitz_nag_NVP.py:286-292,372-374;
SV_dense.py:251-254,330-332): `(lf_sample - c)^2` over every variable and `(theta - theta*)^2` over the
theta-flow variables, each with its own Adamax slots, both gradients taken at the pre-update values.
"""
from __future__ import annotations

import math
import os
import time
from datetime import datetime
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import feed
from .config import OBJ_ELBO, OBJ_PATH_SQ, NMAConfig, fhn_config, lvb_config, lvr_config, param_layout, sv_config
from .engine import NMAEngine
from .theta_flow import ThetaFlow, prior_log_prob, prior_tensors
from .trainer import glorot_blob


class _ModelVISSM:
    """Shared mechanics; the subclasses supply the configuration, the base arrays and the path layout."""

    grad_clip = 2.5e11
    pretrain_path_target = 0.0
    theta_star: Sequence[float] = ()
    theta_pos_index: Sequence[bool] = ()
    has_obs_term = True
    finite_term = 0              # which per-row term the pre-training restart looks at (0 = sde, 2 = lf_log_prob)
    pretrain_theta_optimiser = True      # the scripts' second pre-train optimiser on (theta - theta*)^2 is actually run

    def _common(self, theta_dist: ThetaFlow, priors, p, kernel_len, batch_dims, network_dims, target_dims, no_flows,
                feat_window, learn_rate, pre_train, device, seed, early_stopping):
        if len(set(network_dims)) != 1:
            raise ValueError("all network_dims must be equal (the reference only ever uses [50]*n)")
        self.theta_dist = theta_dist
        self.priors = list(priors)
        self.p = int(p)
        self.kernel_len = int(kernel_len)
        self.batch_dims = int(batch_dims)
        self.network_dims = list(network_dims)
        self.target_dims = int(target_dims)
        self.no_flows = int(no_flows)
        self.feat_window = int(feat_window)
        self.learn_rate = learn_rate
        self.pre_train = pre_train
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.seed = seed
        self.early_stopping = early_stopping
        self.eng: Optional[NMAEngine] = None
        self._scalars_host = {}
        self.pre_train_count = 0

    # supplied by the subclass -------------------------------------------------
    cfg: NMAConfig

    def _base_arrays(self):
        raise NotImplementedError

    def _paths_from_lf(self, lf: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    # ------------------------------------------------------------------
    def build_flow(self) -> None:
        cfg = self.cfg
        self.eng = NMAEngine(cfg, self.device)
        self.eng.set_series(self._base_arrays())
        g = torch.Generator().manual_seed(self.seed)
        _, self.n_nma = param_layout(cfg)
        self.blob = torch.cat([glorot_blob(cfg, g), self.theta_dist.init_values(g)]).to(self.device)
        self.n_total = self.blob.numel()
        zeros = lambda: (torch.zeros_like(self.blob), torch.zeros_like(self.blob))
        self.slots = {"pre_path": zeros(), "pre_theta": zeros(), "main": zeros()}
        self.grad = torch.zeros_like(self.blob)
        self.grad2 = torch.zeros_like(self.blob)
        self.theta_leaf = self.blob[self.n_nma:].detach().requires_grad_(True)
        self.theta_dist.bind(self.theta_leaf)
        self.out = self.eng.alloc_outputs(self.p)
        self.out["grad_params"] = self.grad[:self.n_nma]
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(self.seed)
        self.idx_dev = torch.empty(self.p, dtype=torch.int64, device=self.device)
        self.prior_t = prior_tensors(self.priors, self.device)
        self.theta_star_t = torch.tensor(list(self.theta_star), dtype=torch.float32, device=self.device)
        # main iterations are one nma_train_step call, captured into a CUDA graph after the first eager one
        # (NMA_HOST_THETA=1 keeps the host autograd theta posterior; NMA_FACADE_GRAPH=0 launches eagerly)
        self.eng.set_theta_flow(self.theta_dist, self.priors)
        self.eng.set_seed(self.seed, 0)
        self.scalars_dev = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._theta_last = torch.zeros(self.p, cfg.dtheta, dtype=torch.float32, device=self.device)
        self._host_theta = os.environ.get("NMA_HOST_THETA") == "1"
        self._use_graph = os.environ.get("NMA_FACADE_GRAPH", "1") != "0"
        self._graph, self._main_seen = None, False

    # ------------------------------------------------------------------
    def _iteration(self, batch_select: np.ndarray, pre_train: bool) -> bool:
        """One sess.run of the reference.  Returns False when the step met a non-finite transition density
        (the FHN script restarts its pre-train counter on that, fitz_nag_NVP.py:378-381)."""
        cfg = self.cfg
        self.idx_dev.copy_(torch.from_numpy(np.ascontiguousarray(batch_select, dtype=np.int64)))
        z0 = self.theta_dist.base_sample(self.p, self.gen, self.device)
        theta, logq_theta = self.theta_dist.sample_and_log_prob(z0)
        eps = torch.randn(self.p, cfg.L0, device=self.device, generator=self.gen)
        if pre_train:
            out = self.eng.elbo_fwd_bwd(self.blob[:self.n_nma], eps, theta.detach().contiguous(), self.idx_dev,
                                        objective=OBJ_PATH_SQ, path_target=self.pretrain_path_target, out=self.out)
            # t1: (lf_sample - c)^2 reaches the theta-flow variables through the theta-bias of every flow layer
            self.theta_leaf.grad = None
            (out["grad_theta"] * theta).sum().backward(retain_graph=self.pretrain_theta_optimiser)
            self.grad[self.n_nma:].copy_(self.theta_leaf.grad)
            if self.pretrain_theta_optimiser:
                # t2: (theta - theta*)^2, theta-flow variables only
                self.theta_leaf.grad = None
                ((theta - self.theta_star_t) ** 2).sum().backward()
                self.grad2.zero_()
                self.grad2[self.n_nma:].copy_(self.theta_leaf.grad)
            m, v = self.slots["pre_path"]
            self.eng.adamax_step(self.blob, self.grad, m, v, 1e-3, 0.9, clip=0.0)
            if self.pretrain_theta_optimiser:
                m, v = self.slots["pre_theta"]
                tail = slice(self.n_nma, self.n_total)
                self.eng.adamax_step(self.blob[tail], self.grad2[tail], m[tail], v[tail], 1e-3, 0.9, clip=0.0)
            if self.theta_dist.tf_mask_grad:
                self.theta_dist.constrain()
            return bool(torch.isfinite(out["terms"][:, self.finite_term]).all().item())
        if not self._host_theta:
            self._main_iteration()
            return True
        out = self.eng.elbo_fwd_bwd(self.blob[:self.n_nma], eps, theta.detach().contiguous(), self.idx_dev,
                                    objective=OBJ_ELBO, out=self.out)
        prior = self._prior_lp(theta)
        host_loss = (out["grad_theta"] * theta).sum() - (prior - logq_theta).sum()
        self.theta_leaf.grad = None
        host_loss.backward()
        self.grad[self.n_nma:].copy_(self.theta_leaf.grad)
        m, v = self.slots["main"]
        norm = self.eng.adamax_step(self.blob, self.grad, m, v, self.learn_rate, 0.95, clip=self.grad_clip)
        t = out["terms"]
        scale = float(cfg.scale)
        obs = t[:, 1] if self.has_obs_term else torch.zeros_like(t[:, 1])
        elbo = scale * (t[:, 0] - t[:, 2] + obs) + prior.detach() - logq_theta.detach()
        self._scalars_host = {"loss/ELBO": elbo.mean(), "loss/SDE_log_prob": scale * t[:, 0].mean(),
                        "loss/theta_log_prob": logq_theta.detach().mean(),
                        "loss/path_log_prob": scale * t[:, 2].mean(), "optimize/global_norm": norm.clone()[0]}
        if self.has_obs_term:
            self._scalars_host["loss/obs_log_prob"] = scale * t[:, 1].mean()
        self._theta_last = theta.detach()
        if self.theta_dist.tf_mask_grad:
            self.theta_dist.constrain()
        return True

    def _prior_lp(self, theta: torch.Tensor) -> torch.Tensor:
        """log prior(theta) of the host-composed iteration: the diagonal Gaussian of AR.py:178-182."""
        return prior_log_prob(theta, self.prior_t)

    def _main_body(self) -> None:
        m, v = self.slots["main"]
        self.eng.train_step(self.blob, self.grad, m, v, self.idx_dev, self.scalars_dev, objective=OBJ_ELBO, prior_on=True,
                            obs_in_elbo=self.has_obs_term, lr=self.learn_rate, beta1=0.95, clip=self.grad_clip,
                            theta_out=self._theta_last)

    def _main_iteration(self) -> None:
        """One ELBO iteration: a single nma_train_step call, replayed from a CUDA graph after the first eager one."""
        if not self._use_graph:
            return self._main_body()
        if self._graph is None:
            if not self._main_seen:
                self._main_seen = True
                return self._main_body()
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                self._main_body()
        self._graph.replay()

    @property
    def scalars(self) -> dict:
        return self.read_scalars()

    def read_scalars(self) -> dict:
        """The summaries of the last iteration - ONE device-to-host copy."""
        if self._host_theta:
            return {k: float(v) for k, v in self._scalars_host.items()}
        vals = self.scalars_dev.cpu().tolist()
        out = {"loss/ELBO": vals[0], "loss/SDE_log_prob": vals[1], "loss/theta_log_prob": vals[2],
               "loss/path_log_prob": vals[4], "optimize/global_norm": vals[5]}
        if self.has_obs_term:
            out["loss/obs_log_prob"] = vals[3]
        return out

    def _draw(self) -> np.ndarray:
        replace_bool = bool(self.batch_dims * self.p >= self.target_dims)
        return np.random.choice(np.arange(0, self.target_dims, self.batch_dims), size=self.p, replace=replace_bool)

    def _pretrain_done(self, run: int, finite: bool) -> bool:
        raise NotImplementedError

    def train(self, tensorboard_path, save_path, log_every: int = 1):
        for d in (tensorboard_path, os.path.dirname(save_path)):
            if d and not os.path.exists(d):
                os.makedirs(d)
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter('%s/%s' % (tensorboard_path, datetime.now().strftime("%d:%m:%y-%H:%M:%S")))
        except Exception:       # tensorboard not installed: train without event files
            writer = None
        run = 0
        print("Training model...")
        t_start = time.time()
        while True:
            batch_select = self._draw()
            if self.pre_train:
                if run == 0:
                    print("Pre-training...")
                finite = self._iteration(batch_select, pre_train=True)
                if self._pretrain_done(run, finite):
                    self.pre_train = False
                    run = 0
                    print("Finished pre-training...")
                    self.save(save_path)
            else:
                self._iteration(batch_select, pre_train=False)
                if writer is not None and run % log_every == 0:
                    for tag, val in self.read_scalars().items():
                        writer.add_scalar(tag, float(val), run)
                    th = self._theta_last
                    for i, pos in enumerate(self.theta_pos_index):
                        writer.add_histogram("parameters/%d" % i, (th[:, i].exp() if pos else th[:, i]).cpu(), run)
                if run >= self.early_stopping:
                    break
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
        from .vi_ssm import restore_theta_perms
        restore_theta_perms(self, ck)
        print("Model restored")

    def sample_paths(self, temp_index: int) -> torch.Tensor:
        """lf_sample [p, 2, batch_dims + 1] of p posterior samples of the subsequence starting at `temp_index`."""
        cfg = self.cfg
        idx = torch.full((self.p,), int(temp_index), dtype=torch.int64, device=self.device)
        with torch.no_grad():
            z0 = self.theta_dist.base_sample(self.p, self.gen, self.device)
            theta, _ = self.theta_dist.sample_and_log_prob(z0)
        eps = torch.randn(self.p, cfg.L0, device=self.device, generator=self.gen)
        _, lf = self.eng.forward_paths(self.blob[:self.n_nma], eps, theta.contiguous(), idx)
        return self._paths_from_lf(lf, idx)

    def save_paths(self, PATH_obs):
        """fitz_nag_NVP.py:409-448 / SV_dense.py:363-402: every row evaluates the same subsequence, the windows are
        concatenated along time, the [p, 2, target_dims] tensor is written as [p, 2*target_dims]."""
        path_store = []
        for index_temp in np.arange(0, self.target_dims, self.batch_dims):
            path_store.append(self.sample_paths(int(index_temp))[:, :, 1:].cpu().numpy())
        paths = np.concatenate(path_store, axis=2)
        with open(PATH_obs, 'w') as f:
            np.savetxt(f, np.reshape(paths, (self.p, -1)))
        return paths


class FHN_VI_SSM(_ModelVISSM):
    """fitz_nag_NVP.py:158-448 (class VI_SSM there)."""

    grad_clip = 2.5e11                                                   # fitz_nag_NVP.py:321
    pretrain_path_target = 0.0                                           # :288-289
    theta_star = (math.log(2.0), 1.0, 1.5, math.log(0.5), math.log(0.3))  # :291-292
    theta_pos_index = (True, False, False, True, True)                   # :307
    has_obs_term = True

    def __init__(self, obs, obs_bin, time_till, x0, theta_dist: ThetaFlow, priors, dt, T, p, kernel_len, batch_dims,
                 network_dims, target_dims, no_flows, feat_window, learn_rate=1e-3, pre_train=True, train_paths=True,
                 device: Optional[torch.device] = None, seed: int = 1, early_stopping=1e99):
        self._common(theta_dist, priors, p, kernel_len, batch_dims, network_dims, target_dims, no_flows, feat_window,
                     learn_rate, pre_train, device, seed, early_stopping)
        self.flow_dims = 2
        self.dt, self.T = float(dt), float(T)
        self.kernel_ext = self.kernel_len * self.no_flows + self.flow_dims * self.batch_dims + 2
        self._series = (np.asarray(obs), np.asarray(obs_bin), np.asarray(time_till))
        self.cfg = fhn_config(p=self.p, K=self.kernel_len, B=self.batch_dims, F=self.no_flows,
                              H=len(self.network_dims) - 2, feat_window=self.feat_window,
                              target_dims=self.target_dims, dt=self.dt)
        self.cfg.x0 = (float(x0[0]), float(x0[1]))

    def _base_arrays(self):
        obs, obs_bin, tt = self._series
        return feed.fhn_base_arrays(obs, obs_bin, tt, self.dt, self.T, self.target_dims, self.no_flows,
                                    self.kernel_len, self.feat_window)

    def _paths_from_lf(self, lf, idx):
        # lf_sample = transpose(reshape(lf_sample_init, [p, -1, 2]), [0, 2, 1])  (fitz_nag_NVP.py:283-284)
        return lf.reshape(self.p, -1, 2).transpose(1, 2)

    def _pretrain_done(self, run, finite):
        self.pre_train_count = self.pre_train_count + 1 if finite else 0      # fitz_nag_NVP.py:378-381
        return self.pre_train_count == 500


class LVR_VI_SSM(_ModelVISSM):
    """lotka_volterra_partial.py:160-463 (class VI_SSM there): Lotka-Volterra with a learned theta posterior, on the
    dat/LV_*.txt series the reference ships.  The flow kernels are those of the fixed-theta script; the ELBO branch
    (NMA_MODEL_LVR) is checked on the CPU (tests/test_lvr_formulas.py) but has not run on hardware yet - the library refuses
    the model unless NMA_UNVERIFIED=1 is set."""

    grad_clip = 1e9                                                      # lotka_volterra_partial.py:335
    pretrain_path_target = 75.0                                          # :299-300
    theta_star = (math.log(0.5), math.log(0.0025), math.log(0.3))         # :302-303
    theta_pos_index = (True, True, True)                                 # :306
    has_obs_term = True
    finite_term = 2                                                      # the script watches lf_log_prob (:390-396)

    def __init__(self, obs, obs_bin, time_till, x0, theta_dist: ThetaFlow, priors, dt, T, p, kernel_len, batch_dims,
                 network_dims, target_dims, no_flows, feat_window, learn_rate=1e-3, pre_train=True,
                 device: Optional[torch.device] = None, seed: int = 1, early_stopping=1e99):
        self._common(theta_dist, priors, p, kernel_len, batch_dims, network_dims, target_dims, no_flows, feat_window,
                     learn_rate, pre_train, device, seed, early_stopping)
        self.flow_dims = 2
        self.dt, self.T = float(dt), float(T)
        self.kernel_ext = self.kernel_len * self.no_flows + self.flow_dims * self.batch_dims + 2
        self._series = (np.asarray(obs), np.asarray(obs_bin), np.asarray(time_till))
        self.cfg = lvr_config(p=self.p, K=self.kernel_len, B=self.batch_dims, F=self.no_flows,
                              H=len(self.network_dims) - 2, feat_window=self.feat_window,
                              target_dims=self.target_dims, dt=self.dt, x0=(float(x0[0]), float(x0[1])))

    def _base_arrays(self):
        obs, obs_bin, tt = self._series
        return feed.lvr_base_arrays(obs, obs_bin, tt, self.dt, self.T, self.target_dims, self.no_flows,
                                    self.kernel_len, self.feat_window)

    def _paths_from_lf(self, lf, idx):
        # the library returns the state (softplus * mask + shift applied), interleaved as the flow lays it out
        return lf.reshape(self.p, -1, 2).transpose(1, 2)

    def _pretrain_done(self, run, finite):
        self.pre_train_count = self.pre_train_count + 1 if finite else 0      # lotka_volterra_partial.py:392-398
        return self.pre_train_count == 1000


class LVB_VI_SSM(_ModelVISSM):
    """lotka_volterra_partial_batch.py:190-675 (class VI_SSM there): the fixed-theta script's flow, feed and observation
    model with a LEARNED theta - posterior and prior both end in a Softplus bijector (:358-365,741) -, the plain bivariate
    transition density (:339-343) and p_val windows per iteration that tile p_val concatenated series (:493-494)."""

    grad_clip = 1e9                                                      # lotka_volterra_partial_batch.py:461
    pretrain_path_target = 75.0                                          # :416-417
    theta_pos_index = ()                                                 # the script's theta histograms are commented out (:437-445)
    has_obs_term = True
    finite_term = 2                                                      # the script watches lf_log_prob (:527-531)
    pretrain_theta_optimiser = False                                     # `t2` is built (:419-420) but never run (:525)

    def __init__(self, obs, obs_bin, time_till, x0_mean, x0_std, theta_dist: ThetaFlow, priors, dt, T, p_val, kernel_len,
                 batch_dims, network_dims, target_dims, no_flows, feat_window, learn_rate=1e-3, pre_train=True,
                 device: Optional[torch.device] = None, seed: int = 1, early_stopping=1e99):
        if not getattr(theta_dist, "softplus_out", False):
            raise ValueError("the posterior of this script ends in tfb.Softplus (:741): build ThetaFlow(..., softplus_out=True)")
        x0_std = np.asarray(x0_std, dtype=np.float64)
        if not np.all(x0_std == x0_std[0]):
            raise ValueError("x0_std must be the same for both components")
        self._common(theta_dist, priors, p_val, kernel_len, batch_dims, network_dims, target_dims, no_flows, feat_window,
                     learn_rate, pre_train, device, seed, early_stopping)
        self.p_val = self.p
        self.flow_dims = 2
        self.dt, self.T = float(dt), float(T)
        self.kernel_ext = self.kernel_len * self.no_flows + self.flow_dims * self.batch_dims + 2
        self._series = (np.asarray(obs), np.asarray(obs_bin), np.asarray(time_till))
        self.cfg = lvb_config(p=self.p_val, K=self.kernel_len, B=self.batch_dims, F=self.no_flows,
                              H=len(self.network_dims) - 2, feat_window=self.feat_window,
                              target_dims=self.target_dims, dt=self.dt, x0=np.asarray(x0_mean, dtype=np.float64))
        self.cfg.obs_std = float(x0_std[0])          # this model's only use of the field: the scale of p(x0)

    def _base_arrays(self):
        obs, obs_bin, tt = self._series
        return feed.lv_base_arrays(obs, obs_bin, tt, self.dt, self.T, self.target_dims, self.no_flows, self.kernel_len,
                                   self.feat_window, p_val=self.p_val)

    def _draw(self) -> np.ndarray:
        return feed.sample_indices_lv(self.target_dims, self.batch_dims, self.p_val)       # :493-494, never with replacement

    def _paths_from_lf(self, lf, idx):
        return lf.reshape(self.p, -1, 2).transpose(1, 2)

    def _pretrain_done(self, run, finite):
        self.pre_train_count = self.pre_train_count + 1 if finite else 0      # :527-534
        return self.pre_train_count == 1000

    def _prior_lp(self, theta):
        """TransformedDistribution(MultivariateNormalDiag(mean, scale), Softplus).log_prob(theta) (:358-365)."""
        mean, scale = self.prior_t
        u = theta + torch.log(-torch.expm1(-theta))
        return (-0.5 * ((u - mean) / scale) ** 2 - 0.5 * math.log(2 * math.pi) - torch.log(scale)
                - torch.log(-torch.expm1(-theta))).sum(dim=1)

    def train(self, tensorboard_path, save_path, series_idx=None, num_epochs: int = 3000, log_every: int = 1):
        """:471-600: a fixed number of epochs; pre-training until 1000 consecutive steps with a finite lf_log_prob."""
        self.early_stopping = 1e99
        series_idx_str = 'series_' + str(series_idx) + '_' if series_idx is not None else ''
        for d in (tensorboard_path, os.path.dirname(save_path)):
            if d and not os.path.exists(d):
                os.makedirs(d)
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter('%s/%s' % (tensorboard_path, series_idx_str + datetime.now().strftime("%d:%m:%y-%H:%M:%S")))
        except Exception:
            writer = None
        run = 0
        for epoch in range(int(num_epochs)):
            batch_select = self._draw()
            if self.pre_train:
                if run == 0:
                    print("Initialising paths and parameters...")
                finite = self._iteration(batch_select, pre_train=True)
                if self._pretrain_done(run, finite):
                    self.pre_train = False
                    print("Finished pre-training...")
                    run = 0
            else:
                self._iteration(batch_select, pre_train=False)
                if writer is not None and run % log_every == 0:
                    sc = self.read_scalars()
                    for tag, key in (("loss/ELBO", "loss/ELBO"), ("loss/SDE_log_prob p(x)", "loss/SDE_log_prob"),
                                     ("loss/theta_log_prob q(theta)", "loss/theta_log_prob"),
                                     ("loss/obs_log_prob p(y|x)", "loss/obs_log_prob"),
                                     ("loss/path_log_prob q(x)", "loss/path_log_prob"),
                                     ("optimize/global_norm", "optimize/global_norm")):
                        writer.add_scalar(tag, float(sc[key]), run)
            if run % 1000 == 0:
                self.save(save_path)
            run += 1
        if writer is not None:
            writer.close()

    def save_paths(self, PATH_obs):
        """:605-675: windows at 0, batch_dims, ... < batch_dims * p_val, concatenated along time, written [p_val, 2T]."""
        path_store = []
        for index_temp in np.arange(0, self.batch_dims * self.p_val, self.batch_dims):
            path_store.append(self.sample_paths(int(index_temp))[:, :, 1:].cpu().numpy())
        paths = np.concatenate(path_store, axis=2)
        with open(PATH_obs, 'w') as f:
            np.savetxt(f, np.reshape(paths, (self.p_val, -1)))
        return paths


class SV_VI_SSM(_ModelVISSM):
    """SV_dense.py:139-402 (class VI_SSM there)."""

    grad_clip = 1e7                                                      # SV_dense.py:283
    pretrain_path_target = -7.0                                          # :251-252
    theta_star = (0.001, -0.6, math.log(0.08), math.log(0.5))             # :253-254
    theta_pos_index = (False, False, True, True)                         # :257
    has_obs_term = False

    def __init__(self, obs, x0, theta_dist: ThetaFlow, priors, dt, T, p, kernel_len, batch_dims, network_dims,
                 target_dims, no_flows, feat_window, learn_rate=1e-3, pre_train=False,
                 device: Optional[torch.device] = None, seed: int = 1, early_stopping=1e99, exact_var: bool = True):
        self._common(theta_dist, priors, p, kernel_len, batch_dims, network_dims, target_dims, no_flows, feat_window,
                     learn_rate, pre_train, device, seed, early_stopping)
        self.dt, self.T, self.x0 = float(dt), float(T), float(x0)
        self.kernel_ext = self.kernel_len * self.no_flows + self.batch_dims + 1
        self.obs = np.asarray(obs)
        last = ((self.target_dims - 1) // self.batch_dims) * self.batch_dims
        if last + self.batch_dims + 1 > self.obs.shape[0]:
            # the reference's np.concatenate of obs[index : index + batch_dims + 1] slices (SV_dense.py:327-328) raises
            # on the ragged last window; say why instead
            raise ValueError("the last subsequence (start %d, %d + 1 prices) runs past the series (%d prices): "
                             "choose batch_dims dividing target_dims" % (last, self.batch_dims, self.obs.shape[0]))
        self.exact_var = exact_var
        self.cfg = sv_config(p=self.p, K=self.kernel_len, B=self.batch_dims, F=self.no_flows,
                             H=len(self.network_dims) - 2, feat_window=self.feat_window,
                             target_dims=self.target_dims, dt=self.dt, x0=self.x0)

    def _base_arrays(self):
        var_fn = None
        if self.exact_var and self.obs.dtype == np.float32:
            # A14 on the device: nma_rolling_var is bit-exact with the script's float32 np.var loop (SV_dense.py:159-170)
            from .engine import rolling_var

            def var_fn(x, K):
                x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.device)
                return rolling_var(x, K).cpu().numpy()
        return feed.sv_base_arrays(self.obs, self.dt, self.T, self.no_flows, self.kernel_len, self.feat_window,
                                   exact_var=self.exact_var, var_fn=var_fn)

    def build_flow(self) -> None:
        super().build_flow()
        self._obs_dev = torch.from_numpy(self.obs.astype(np.float32)).to(self.device)

    def _paths_from_lf(self, lf, idx):
        # lf_sample = [dim_one, lf_sample_temp * mask + shift]  (SV_dense.py:243-245): the observed price and the
        # latent log-volatility whose very first value is pinned to x0
        B1 = self.batch_dims + 1
        t = idx[:, None] + torch.arange(B1, device=self.device)[None, :]
        dim_one = self._obs_dev[t.clamp(max=self._obs_dev.numel() - 1)]
        first = (t == 0)
        latent = torch.where(first, torch.full_like(lf[:, :B1], self.x0), lf[:, :B1])
        return torch.stack([dim_one, latent], dim=1)

    def _pretrain_done(self, run, finite):
        return run == 1000                                               # SV_dense.py:335-338


class LV_VI_SSM:
    """lotka_volterra_partial_batch_fix_theta.py:181-613 (class VI_SSM there): the Lotka-Volterra model with the
    parameters fixed at `priors` (theta is a constant tiled over the rows, :190 - there is no theta posterior, no
    prior term and no second optimiser), p_val subsequences per iteration drawn without replacement from
    arange(0, target_dims * p_val, batch_dims) (:478-479), pre-training on (lf_sample - 75)^2 until
    `pre_train_epochs` consecutive finite steps (:505-516), then the ELBO with clip_by_global_norm(1e9) (:450-458).

    The committed script runs p_val = 1 with batch_dims = target_dims = 151: every iteration evaluates the one
    subsequence that is the whole series.  The mask/shift pin of the first state is reproduced for that case
    (`mask_vals = zeros((2, p_val)) ++ ones`, :216-219, coincides with "index 0 only" when p_val = 1)."""

    grad_clip = 1e9
    pretrain_path_target = 75.0

    def __init__(self, obs, obs_bin, time_till, x0_mean, x0_std, priors, dt, T, p_val, kernel_len, batch_dims,
                 network_dims, target_dims, no_flows, feat_window, learn_rate=1e-3, pre_train=True,
                 device: Optional[torch.device] = None, seed: int = 1):
        from .config import lv_config
        if len(set(network_dims)) != 1:
            raise ValueError("all network_dims must be equal (the reference only ever uses [50]*n)")
        if int(p_val) != 1:
            raise ValueError("p_val > 1 changes which states are pinned to x0 (mask_vals, :216-219); only the "
                             "script's p_val = 1 is reproduced")
        x0_std = np.asarray(x0_std, dtype=np.float64)
        if not np.all(x0_std == x0_std[0]):
            raise ValueError("x0_std must be the same for both components")
        self.p_val = int(p_val)
        self.priors = np.asarray(priors, dtype=np.float64)
        self.kernel_len, self.batch_dims = int(kernel_len), int(batch_dims)
        self.network_dims, self.no_flows = list(network_dims), int(no_flows)
        self.target_dims, self.feat_window = int(target_dims), int(feat_window)
        self.dt, self.T = float(dt), float(T)
        self.learn_rate, self.pre_train = learn_rate, pre_train
        self.flow_dims = 2
        self.kernel_ext = self.kernel_len * self.no_flows + self.flow_dims * self.batch_dims + 2
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.seed = seed
        self._series = (np.asarray(obs), np.asarray(obs_bin), np.asarray(time_till))
        self.cfg = lv_config(p=self.p_val, K=self.kernel_len, B=self.batch_dims, F=self.no_flows,
                             H=len(self.network_dims) - 2, feat_window=self.feat_window,
                             target_dims=self.target_dims, dt=self.dt, x0=np.asarray(x0_mean, dtype=np.float64))
        self.cfg.obs_std = float(x0_std[0])          # this model's only use of the field: the scale of p(x0)
        self.eng: Optional[NMAEngine] = None

    def build_flow(self) -> None:
        cfg = self.cfg
        self.eng = NMAEngine(cfg, self.device)
        obs, obs_bin, tt = self._series
        self.eng.set_series(feed.lv_base_arrays(obs, obs_bin, tt, self.dt, self.T, self.target_dims, self.no_flows,
                                                self.kernel_len, self.feat_window, p_val=self.p_val))
        g = torch.Generator().manual_seed(self.seed)
        self.blob = glorot_blob(cfg, g).to(self.device)
        self.n_nma = self.n_total = self.blob.numel()
        zeros = lambda: (torch.zeros_like(self.blob), torch.zeros_like(self.blob))
        self.slots = {"pre_path": zeros(), "main": zeros()}
        self.grad = torch.zeros_like(self.blob)
        self.out = self.eng.alloc_outputs(self.p_val)
        self.out["grad_params"] = self.grad
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(self.seed)
        self.idx_dev = torch.empty(self.p_val, dtype=torch.int64, device=self.device)
        self.theta = torch.tensor(self.priors, dtype=torch.float32, device=self.device).repeat(self.p_val, 1).contiguous()
        # ELBO iterations: one nma_train_step call (in-library noise, constant theta), replayed from a CUDA graph
        self.eng.set_fixed_theta([float(v) for v in self.priors])
        self.eng.set_seed(self.seed, 0)
        self.scalars_dev = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._lf_dev = torch.zeros(self.p_val, cfg.L(cfg.F), dtype=torch.float32, device=self.device)
        self._use_graph = os.environ.get("NMA_FACADE_GRAPH", "1") != "0"
        self._graph, self._main_seen = None, False

    def _iteration(self, batch_select: np.ndarray, pre_train: bool) -> bool:
        cfg = self.cfg
        self.idx_dev.copy_(torch.from_numpy(np.ascontiguousarray(batch_select, dtype=np.int64)))
        if pre_train:
            eps = torch.randn(self.p_val, cfg.L0, device=self.device, generator=self.gen)
            out = self.eng.elbo_fwd_bwd(self.blob, eps, self.theta, self.idx_dev, objective=OBJ_PATH_SQ,
                                        path_target=self.pretrain_path_target, out=self.out)
            m, v = self.slots["pre_path"]
            self.eng.adamax_step(self.blob, self.grad, m, v, 1e-3, 0.9, clip=0.0)
            self.lf_sample = out["lf"].reshape(self.p_val, -1, 2).transpose(1, 2)
            return bool(torch.isfinite(out["terms"][:, 2]).all().item())      # `test = lf_log_prob`, :507-510
        self._main_iteration()
        self.lf_sample = self._lf_dev.reshape(self.p_val, -1, 2).transpose(1, 2)
        return True

    def _prior_lp(self, theta: torch.Tensor) -> torch.Tensor:
        """log prior(theta) of the host-composed iteration: the diagonal Gaussian of AR.py:178-182."""
        return prior_log_prob(theta, self.prior_t)

    def _main_body(self) -> None:
        m, v = self.slots["main"]
        self.eng.train_step(self.blob, self.grad, m, v, self.idx_dev, self.scalars_dev, objective=OBJ_ELBO, prior_on=False,
                            lr=self.learn_rate, beta1=0.95, clip=self.grad_clip, lf_out=self._lf_dev)

    def _main_iteration(self) -> None:
        if not self._use_graph:
            return self._main_body()
        if self._graph is None:
            if not self._main_seen:
                self._main_seen = True
                return self._main_body()
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                self._main_body()
        self._graph.replay()

    @property
    def scalars(self) -> dict:
        """The script's summaries (:436-448) of the last ELBO iteration - one device-to-host copy."""
        v = self.scalars_dev.cpu().tolist()
        return {"loss/NELBO": -v[0], "loss/ELBO": v[0], "loss/SDE_log_prob p(x)": v[1], "loss/obs_log_prob p(y|x)": v[3],
                "loss/path_log_prob q(x)": v[4], "optimize/global_norm": v[5]}

    def _draw(self) -> np.ndarray:
        return feed.sample_indices_lv(self.target_dims, self.batch_dims, self.p_val)

    def train(self, tensorboard_path, save_path, num_epochs, pre_train_epochs, series_idx=None):
        series_idx_str = 'series_' + str(series_idx) + '_' if series_idx is not None else ''
        for d in (tensorboard_path, os.path.dirname(save_path)):
            if d and not os.path.exists(d):
                os.makedirs(d)
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter('%s/%s' % (tensorboard_path,
                                              series_idx_str + datetime.now().strftime("%d:%m:%y-%H:%M:%S")))
        except Exception:
            writer = None
        run, pre_train_count = 0, 0
        t_start = time.time()
        self.epoch_times = []
        for epoch in range(num_epochs):
            batch_select = self._draw()
            if self.pre_train:
                if run == 0:
                    print("Initialising paths and parameters...")
                pre_train_count = pre_train_count + 1 if self._iteration(batch_select, pre_train=True) else 0
                if pre_train_count == pre_train_epochs:
                    self.pre_train = False
                    print("Finished pre-training...")
                    run = 0
            else:
                self._iteration(batch_select, pre_train=False)
                if writer is not None:
                    for tag, val in self.scalars.items():
                        writer.add_scalar(tag, float(val), run)
                    writer.add_scalar("elapsed_time", time.time() - t_start, epoch)
            if run % 1000 == 0:
                self.save(save_path)
            run += 1
            self.epoch_times.append(time.time() - t_start)
        if writer is not None:
            writer.close()
        return self.lf_sample

    def save(self, PATH):
        torch.save({"blob": self.blob.cpu(), "slots": {k: (m.cpu(), v.cpu()) for k, (m, v) in self.slots.items()},
                    "numpy_rng": np.random.get_state()}, PATH)
        print("Model saved")

    def load(self, PATH):
        self.pre_train = False
        ck = torch.load(PATH, weights_only=False)
        self.blob.copy_(ck["blob"].to(self.device))
        for k, (m, v) in ck["slots"].items():
            self.slots[k][0].copy_(m.to(self.device))
            self.slots[k][1].copy_(v.to(self.device))
        print("Model restored")

    def sample_paths(self, temp_index: int) -> torch.Tensor:
        """lf_sample [p_val, 2, batch_dims + 1] (softplus-transformed, first state pinned) of one posterior draw."""
        idx = torch.full((self.p_val,), int(temp_index), dtype=torch.int64, device=self.device)
        eps = torch.randn(self.p_val, self.cfg.L0, device=self.device, generator=self.gen)
        _, lf = self.eng.forward_paths(self.blob, eps, self.theta, idx)
        return lf.reshape(self.p_val, -1, 2).transpose(1, 2)

    def save_paths(self, PATH_obs):
        """:571-613: windows at 0, batch_dims, ... < batch_dims * p_val, concatenated along time, written [p_val, 2T]."""
        path_store = []
        for index_temp in np.arange(0, self.batch_dims * self.p_val, self.batch_dims):
            path_store.append(self.sample_paths(int(index_temp))[:, :, 1:].cpu().numpy())
        paths = np.concatenate(path_store, axis=2)
        with open(PATH_obs, 'w') as f:
            np.savetxt(f, np.reshape(paths, (self.p_val, -1)))
        return paths
