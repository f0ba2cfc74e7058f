"""Device-resident training step for the AR(1) NMA model, single- and multi-GPU.

`ARStepper` is the loop body of `VI_SSM.train` (AR.py:262-310) with everything except the index draw on
the device: the series lives in HBM (generated there for long T by the A12/A13 scans, time-sharded across
ranks with a (no_flows*kernel_len+1 | feat_window-1) halo), theta is sampled by the host-side theta flow
(viforssms_b200.theta_flow), the C-ABI library evaluates ELBO + gradients (nma_elbo_fwd_bwd), gradients are
all-reduced over NCCL when world > 1, and the fused clip + Adamax kernel updates ONE flat blob holding the
NMA parameters followed by the theta-flow parameters (one global norm over everything, AR.py:228-234).
"""
from __future__ import annotations

import math
import queue
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as _lib
from .config import NMAConfig, ar_config, param_layout
from .engine import NMAEngine, scan_ar1, time_till
from .theta_flow import ThetaFlow, prior_log_prob, prior_tensors


def shard_bounds(T: int, B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous time range [t0, t1) owned by `rank`; boundaries are multiples of batch_dims so the
    candidate starts arange(0, T, B) (AR.py:257) partition exactly."""
    nb = (T + B - 1) // B
    lo = (nb * rank) // world
    hi = (nb * (rank + 1)) // world
    return lo * B, min(hi * B, T)


def glorot_blob(cfg: NMAConfig, gen: torch.Generator) -> torch.Tensor:
    """TF defaults for every NMA variable: Glorot-uniform kernels, zero biases, BN gamma=1 beta=0."""
    layout, n = param_layout(cfg)
    flat = torch.zeros(n, dtype=torch.float32)
    for name, (off, shape) in layout.items():
        k = int(np.prod(shape))
        if name.endswith(".w"):
            if len(shape) == 3:
                fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
            else:
                fan_in, fan_out = shape
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            flat[off:off + k] = (torch.rand(k, generator=gen, dtype=torch.float64) * 2 - 1).float() * lim
        elif name.endswith(".gamma"):
            flat[off:off + k] = 1.0
    return flat


def local_base_arrays(obs_ext: torch.Tensor, bin_ext: torch.Tensor, till_ext: torch.Tensor, tt0: float, t0: int,
                      t1: int, P: int, fw: int) -> List[torch.Tensor]:
    """The five padded base arrays of AR.py:135-150 restricted to the shard [t0, t1).

    `*_ext` hold series values for global steps [t0 - P, t1 + fw - 1) (zeros where that leaves [0, T)): the
    left halo of P = no_flows*kernel_len + 1 and the right halo of feat_window - 1 samples.  Local padded
    index j corresponds to the reference's padded index q = t0 + j, so a window starting at local index
    idx - t0 is the reference's window starting at idx, value for value."""
    dev = obs_ext.device
    n = t1 - t0 + P                                   # padded indices q in [t0, t1 + P)
    q = torch.arange(t0, t0 + n + 1, device=dev, dtype=torch.float64)
    obs_pad = obs_ext.to(torch.float64)               # already [t0-P, t1+fw-1) == padded [t0, t1+P+fw-1)
    bin_feats = (q[:n] < P).to(torch.float64)
    time_pad = torch.clamp(q - P, min=0.0)            # zeros(P) ++ arange(T+1), one longer like the reference
    lead = (P + tt0) - q[:n]                          # arange(P+tt0, tt0, -1) for q < P
    time_till_pad = torch.where(q[:n] < P, lead, till_ext[:n].to(torch.float64))
    obs_bin_pad = torch.where(q[:n] < P, torch.zeros_like(lead), bin_ext[:n].to(torch.float64))
    return [obs_pad, bin_feats, time_pad, time_till_pad, obs_bin_pad]


def exchange_halos(fill: torch.Tensor, binary: torch.Tensor, till: torch.Tensor, P: int, fw: int, rank: int,
                   world: int):
    """Neighbour exchange for a time-sharded series (works on any torch.distributed backend).

    Every rank holds its own steps of the three series (hold-filled observations, observation indicator,
    time-till-next-observation).  A window needs P = no_flows*kernel_len + 1 samples to the left of the shard
    (all three series) and feat_window - 1 observations to the right (look-ahead channels, AR.py:136-138):
    rank r sends its last P samples to r+1 and its first fw-1 observations to r-1.  Outside [0, T) the halos
    are the reference's zero padding.  Returns (obs_ext [P+n+fw-1], bin_ext [P+n], till_ext [P+n], time_till[0] of
    the GLOBAL series, which the leading pad counts down from, AR.py:149-150)."""
    import torch.distributed as dist
    dev, dt = fill.device, fill.dtype
    nr = max(fw - 1, 0)
    left = [torch.zeros(P, dtype=dt, device=dev) for _ in range(3)]
    right = torch.zeros(nr, dtype=dt, device=dev)
    tt0 = torch.tensor([float(till[0].item())], dtype=torch.float64, device=dev)
    if world > 1:
        if fill.numel() < P:
            raise ValueError("a time shard must be at least no_flows*kernel_len+1 steps long")
        dist.broadcast(tt0, src=0)
        send_l = [fill[-P:].contiguous(), binary[-P:].contiguous(), till[-P:].contiguous()]
        send_r = fill[:nr].contiguous()
        ops = []
        if rank + 1 < world:
            ops += [dist.P2POp(dist.isend, t, rank + 1) for t in send_l]
            if nr:
                ops.append(dist.P2POp(dist.irecv, right, rank + 1))
        if rank > 0:
            ops += [dist.P2POp(dist.irecv, t, rank - 1) for t in left]
            if nr:
                ops.append(dist.P2POp(dist.isend, send_r, rank - 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    return (torch.cat([left[0], fill, right]), torch.cat([left[1], binary]), torch.cat([left[2], till]),
            float(tt0.item()))


class IndexFeeder:
    """Background draw of subsequence starts with the reference's own call (AR.py:263-265) into pinned
    host buffers: the legacy permutation of 2*10^6 candidates costs tens of ms and must not sit on the
    critical path of a step."""

    def __init__(self, cand: np.ndarray, rows: int, replace: bool, seed: int, offset: int, depth: int = 4,
                 pinned: bool = True):
        self.cand, self.rows, self.replace, self.offset = cand, rows, replace, offset
        self.rs = np.random.RandomState(seed)
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)
        self.pinned = pinned
        self._stop = False
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def draw(self) -> np.ndarray:
        return self.rs.choice(self.cand, size=self.rows, replace=self.replace).astype(np.int64)

    def _run(self):
        while not self._stop:
            idx = self.draw() - self.offset
            buf = torch.from_numpy(idx)
            if self.pinned:
                buf = buf.pin_memory()
            while not self._stop:
                try:
                    self.q.put(buf, timeout=0.1)
                    break
                except queue.Full:
                    continue

    def get(self) -> torch.Tensor:
        return self.q.get()

    def close(self):
        self._stop = True


class ARStepper:
    def __init__(self, T: int, rows: int, K: int = 50, B: int = 50, F: int = 3, H: int = 1, fw: int = 10,
                 theta: Sequence[float] = (5.0, 0.5, 3.0), x0: float = 10.0, obs_std: float = 1.0,
                 device: Optional[torch.device] = None, rank: int = 0, world: int = 1, seed: int = 1,
                 lr: float = 1e-3, clip: float = 2.5e8, priors=((0.0, 10.0),) * 3, series=None, impute: int = 1,
                 tensor_cores: Optional[int] = None, device_theta: bool = True, tf_mask_grad: bool = True,
                 unscaled_time_weights: bool = False):
        self.T, self.rows, self.rank, self.world = int(T), int(rows), rank, world
        # device_theta=True (default): the whole iteration is ONE nma_train_step call - in-library Philox noise, theta
        # posterior, ELBO + gradients, per-flow NCCL all-reduce, clip + Adamax, logged scalars; no ATen kernel in the step.
        # False: the host autograd theta posterior + torch.randn (the comparison path of the tests).
        self.device_theta = bool(device_theta)
        self.tf_mask_grad = bool(tf_mask_grad)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.lr, self.clip, self.priors = lr, clip, priors
        self.cfg = ar_config(p=rows, K=K, B=B, F=F, H=H, feat_window=fw, T=T, obs_std=obs_std, x0=x0)
        cfg = self.cfg
        self.t0, self.t1 = shard_bounds(self.T, B, world, rank)
        P = F * K + 1
        self.eng = NMAEngine(cfg, self.device, tensor_cores=tensor_cores)   # None: library default; bit set of nma_set_tensor_cores
        if series is None:
            obs_ext, bin_ext, till_ext, tt0 = self._generate_shard(theta, x0, obs_std, seed, P, fw, impute)
        else:
            obs_ext, bin_ext, till_ext, tt0 = self._slice_series(series, P, fw)
        self.eng.set_series(local_base_arrays(obs_ext, bin_ext, till_ext, tt0, self.t0, self.t1, P, fw))

        # ---- parameters: NMA blob ++ theta-flow blob, Adamax slots, gradient blob ----
        g = torch.Generator().manual_seed(seed)
        nma = glorot_blob(cfg, g)
        # the raw time-index channel reaches T (AR.py:139-140): keep first-layer activations O(1) at init
        layout, n_nma = param_layout(cfg)
        # (unscaled_time_weights=True keeps plain Glorot values there, which is what the reference's initialiser gives)
        for i in range(F):
            if unscaled_time_weights:
                break
            off, shape = layout[f"f{i}.feat0.w"]
            nma[off:off + shape[0] * shape[1]].reshape(shape)[fw + 1, :] *= 10.0 / max(self.T, 1)
        perms = [np.random.RandomState(seed + 17 + k).permutation(3) for k in range(4)]
        self.flow = ThetaFlow(3, 5, 1.5, 0.5, "elu", perms, tf_mask_grad=tf_mask_grad)            # AR.py:378-390
        tf_init = self.flow.init_values(g)
        self.n_nma, self.n_total = n_nma, n_nma + self.flow.n_params
        self.blob = torch.cat([nma, tf_init]).to(self.device)
        self.m = torch.zeros_like(self.blob)
        self.v = torch.zeros_like(self.blob)
        self.grad = torch.zeros_like(self.blob)
        self.theta_leaf = self.blob[n_nma:].detach().requires_grad_(True)
        self.flow.bind(self.theta_leaf)
        self.out = self.eng.alloc_outputs(rows)
        self.out["grad_params"] = self.grad[:n_nma]
        # eps / theta base noise come from torch's default CUDA generator (graph-capture safe)
        torch.cuda.manual_seed(seed * 1000 + rank)
        self.prior_t = prior_tensors(priors, self.device)
        self.eng.set_theta_flow(self.flow, list(priors))
        self.eng.set_seed(seed * 1000 + rank, 0)
        self.scalars = torch.zeros(8, dtype=torch.float32, device=self.device)
        import torch.distributed as dist
        if world > 1 and dist.is_available() and dist.is_initialized():
            # the library's own NCCL communicator: the gradient all-reduce is issued inside nma_train_step /
            # nma_elbo_fwd_bwd, per flow, on the library's side stream
            self.eng.comm_create(rank, world)
        self.graph = None
        self._elbo_static = None
        self.launches_per_step = None

        # ---- index feed ----
        cand = np.arange(self.t0, self.t1, B)
        replace = bool(B * rows >= (self.t1 - self.t0))                # AR.py:264-265 on the local range
        self.feeder = IndexFeeder(cand, rows, replace, seed + rank, self.t0)
        self.idx_dev = torch.empty(rows, dtype=torch.int64, device=self.device)
        self.idx_dev.copy_(self.feeder.get())
        self.h2d_bytes = rows * 8
        self.d2h_bytes = 8
        self.last_elbo = None

    # ------------------------------------------------------------------
    def _generate_shard(self, theta, x0, obs_std, seed, P, fw, impute):
        """A12/A13 on the device, time-sharded: every rank scans its own range; the carries (one affine
        map per rank) are all-gathered so each rank restarts from the exact entering value, then the
        P-left / (fw-1)-right halos are exchanged with the neighbours."""
        import torch.distributed as dist
        dev = self.device
        n = self.t1 - self.t0
        a, b, c = float(theta[1]), float(theta[0]), float(theta[2])
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 7919 + self.rank)
        z = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        start = float(x0)
        if self.world > 1:
            xl = scan_ar1(z, 0.0, a, b, c)
            mine = torch.tensor([float(torch.tensor(a, dtype=torch.float64) ** n), float(xl[-1].item())],
                                dtype=torch.float64, device=dev)
            allc = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(allc, mine)
            for r in range(self.rank):
                A, D = (float(v) for v in allc[r].tolist())
                start = A * start + D
            del xl
        X = scan_ar1(z, start, a, b, c)                     # X[0] = value entering the shard, X[1..n] = steps t0+1..t1
        del z
        obs = X + float(obs_std) * torch.randn(n + 1, dtype=torch.float64, device=dev, generator=g)
        del X
        if n % impute != 0:
            raise ValueError("shard length must be a multiple of impute")
        fill, binary, till = time_till(obs.contiguous(), int(impute))
        del obs
        obs_ext, bin_ext, till_ext, tt0 = exchange_halos(fill, binary, till, P, fw, self.rank, self.world)
        return obs_ext.float(), bin_ext.float(), till_ext.float(), tt0

    def _slice_series(self, series, P, fw):
        """Host series (obs, obs_bin, time_till of length T, float64 numpy) -> this rank's extended slices."""
        obs, obs_bin, tt = (np.asarray(s, dtype=np.float64) for s in series)
        T = obs.shape[0]

        def ext(arr, lo, hi):
            out = np.zeros(hi - lo)
            a, b = max(lo, 0), min(hi, T)
            if b > a:
                out[a - lo:b - lo] = arr[a:b]
            return torch.from_numpy(out).to(self.device)
        lo = self.t0 - P
        return (ext(obs, lo, self.t1 + fw - 1).float(), ext(obs_bin, lo, self.t1).float(), ext(tt, lo, self.t1).float(),
                float(tt[0]))

    # ------------------------------------------------------------------
    def _step(self, idx_dev: torch.Tensor) -> torch.Tensor:
        """One iteration = one nma_train_step call (AR.py:300-301); returns the mean ELBO (a view of `scalars`)."""
        if not self.device_theta:
            return self._step_host_theta(idx_dev)
        self.eng.train_step(self.blob, self.grad, self.m, self.v, idx_dev, self.scalars, objective=0, prior_on=True,
                            tf_mask_grad=self.tf_mask_grad, lr=self.lr, beta1=0.95, clip=self.clip)
        return self.scalars[0]

    def _step_host_theta(self, idx_dev: torch.Tensor, z0: Optional[torch.Tensor] = None,
                         eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The same iteration with the theta posterior as a host autograd module and torch.randn noise (or the injected
        `z0` / `eps`): the comparison path of the tests (tests/test_gpu_lvr_theta.py); ~400 small ATen launches."""
        cfg, rows = self.cfg, self.rows
        if z0 is None:
            z0 = self.flow.base_sample(rows, None, self.device)
        theta, logq_theta = self.flow.sample_and_log_prob(z0)
        if eps is None:
            eps = torch.randn(rows, cfg.L0, device=self.device)
        out = self.eng.elbo_fwd_bwd(self.blob[:self.n_nma], eps, theta.detach().contiguous(), idx_dev, out=self.out)
        # host-side remainder of -sum(ELBO): the theta path (AR.py:178-185)
        tail = prior_log_prob(theta, self.prior_t) - logq_theta
        host_loss = (out["grad_theta"] * theta).sum() - tail.sum()
        self.theta_leaf.grad = None
        host_loss.backward()
        self.grad[self.n_nma:].copy_(self.theta_leaf.grad)
        if self.world > 1:
            self.eng.comm_wait()                                   # the flow sections were all-reduced by the library
            self.eng.comm_allreduce(self.grad[self.n_nma:])
        self.eng.adamax_step(self.blob, self.grad, self.m, self.v, self.lr, 0.95, clip=self.clip)
        if self.tf_mask_grad:
            self.flow.constrain()
        t = out["terms"]
        elbo = (float(cfg.scale) * (t[:, 0] - t[:, 2] + t[:, 1]) + tail.detach()).mean()
        return elbo

    def capture(self) -> None:
        """Capture the whole iteration (theta flow, ELBO + gradients, all-reduce, clip + Adamax) into ONE CUDA
        graph: a step becomes a single launch, which is what makes the p=50 reference configuration and the
        host-fed loop stop being launch-bound.  Inputs (`idx_dev`) and outputs (`_elbo_static`) are static."""
        L = _lib.load()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for it in range(3):
                n0 = L.nma_launch_count()
                self._step(self.idx_dev)
                self.launches_per_step = int(L.nma_launch_count() - n0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # thread_local: the index feeder's thread may pin host memory while this thread captures
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._elbo_static = self._step(self.idx_dev)
        self.graph = g

    def step_resident(self) -> torch.Tensor:
        """One step with the subsequence indices already on the device."""
        if self.graph is not None:
            self.graph.replay()
            self.last_elbo = self._elbo_static
        else:
            self.last_elbo = self._step(self.idx_dev)
        return self.last_elbo

    def step_e2e(self) -> float:
        """One step through host buffers: pinned index batch -> device, ELBO scalar -> host."""
        host_idx = self.feeder.get()
        self.idx_dev.copy_(host_idx, non_blocking=True)
        return float(self.step_resident().item())

    # ------------------------------------------------------------------
    def time_stage(self, stage: int, flow: int, reps: int = 5) -> float:
        """Average device time (ms) of ONE kernel of the last step, re-launched on the workspace it left."""
        L = _lib.load()
        st = torch.cuda.current_stream()
        eps = torch.randn(self.rows, self.cfg.L0, device=self.device)
        scratch = torch.zeros_like(self.grad)

        def launch():
            _lib.check(L.nma_launch_stage(self.eng._h, stage, flow, self.blob.data_ptr(), eps.data_ptr(),
                                          self.idx_dev.data_ptr(), self.rows, scratch.data_ptr(), st.cuda_stream),
                       "nma_launch_stage")
        launch()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            launch()
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def close(self):
        """Teardown in the order NCCL needs: stop feeding, drain the device, destroy the captured graph (it holds the
        library's NCCL kernels when world > 1), then the communicator, then the handle."""
        self.feeder.close()
        torch.cuda.synchronize(self.device)
        if self.graph is not None:
            self.graph.reset()
        self.graph = None
        self._elbo_static = None
        torch.cuda.synchronize(self.device)
        self.eng.comm_destroy()
        self.eng.close()
