"""Thin host wrapper over the C-ABI: torch owns device memory and streams, the library does the work.

`NMAEngine` is what the `VI_SSM` facade (viforssms_b200/vi_ssm.py) drives in place of the reference's
`sess.run([train_step, merged], feed_dict)` (AR.py:300-301).
"""
from __future__ import annotations

import ctypes
from ctypes import c_int64, c_void_p
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import lib as _lib
from .config import NMAConfig, param_layout


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


class _DevView:
    """A borrowed fp32 device buffer of the library, seen through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class NMAEngine:
    def __init__(self, cfg: NMAConfig, device: Optional[torch.device] = None, tensor_cores: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise _lib.NMAError("NMAEngine needs a CUDA device: the NMA ELBO step has no CPU path")
        self.cfg = cfg
        self.device = torch.device(device if device is not None else "cuda", torch.cuda.current_device()) \
            if not isinstance(device, torch.device) else device
        self._lib = _lib.load()
        self._h = c_void_p()
        ccfg = cfg.to_c()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.nma_create(ctypes.byref(ccfg), ctypes.byref(self._h)), "nma_create")
        if tensor_cores is not None:
            # bit 0: K-tap conv (forward, data and weight gradient) on tcgen05; bit 1: feature MLP as well.
            # True selects everything, an int selects exactly those bits (1 = conv only).
            mode = (3 if tensor_cores else 0) if isinstance(tensor_cores, bool) else int(tensor_cores)
            _lib.check(self._lib.nma_set_tensor_cores(self._h, mode), "nma_set_tensor_cores")
        self.n_params = int(self._lib.nma_param_count(self._h))
        self.layout, n = param_layout(cfg)
        if n != self.n_params:
            raise _lib.NMAError(f"host/device parameter layouts disagree: {n} vs {self.n_params}")
        self._series: List[torch.Tensor] = []
        self._scratch = torch.zeros(1024, dtype=torch.float32, device=self.device)
        self._norm = torch.zeros(1, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nma_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def tensor_cores(self) -> bool:
        """True when the conv and its data gradient run on the tcgen05 tensor cores (3xTF32)."""
        return (int(self._lib.nma_get_tensor_cores(self._h)) & 1) == 1

    @property
    def tensor_core_features(self) -> bool:
        """True when the feature MLP runs on the tensor cores as well (nma_tc_feat.cu)."""
        return (int(self._lib.nma_get_tensor_cores(self._h)) & 2) == 2

    @property
    def bf16_split(self) -> bool:
        """True when the conv GEMMs run in the 2-term bf16 split on kind::f16 (mode bit 2 / NMA_TC_BF16=1)."""
        return (int(self._lib.nma_get_tensor_cores(self._h)) & 4) == 4

    def set_tensor_cores(self, mode: int) -> None:
        """Bit set of nma_set_tensor_cores: 1 conv, 2 feature MLP + head backward, 4 bf16 split of the conv GEMMs."""
        _lib.check(self._lib.nma_set_tensor_cores(self._h, int(mode)), "nma_set_tensor_cores")

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.nma_workspace_bytes(self._h))

    def device_layout(self) -> np.ndarray:
        out = (c_int64 * (self.cfg.F * 32))()
        _lib.check(self._lib.nma_param_layout(self._h, out, self.cfg.F * 32), "nma_param_layout")
        return np.ctypeslib.as_array(out).reshape(self.cfg.F, 32).copy()

    # ------------------------------------------------------------------
    def set_series(self, arrays: Sequence) -> None:
        """Padded base arrays (float64 numpy as the reference builds them, or fp32 tensors) -> device fp32.

        The float64 -> float32 conversion is the feed_dict cast of the reference (AR.py:300-301)."""
        dev = []
        for a in arrays:
            if isinstance(a, torch.Tensor):
                t = a.to(device=self.device, dtype=torch.float32).contiguous()
            else:
                t = torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(np.float32))).to(self.device)
            dev.append(t.reshape(-1))
        self._series = dev
        n = len(dev)
        ptrs = (c_void_p * n)(*[t.data_ptr() for t in dev])
        lens = (c_int64 * n)(*[t.numel() for t in dev])
        _lib.check(self._lib.nma_set_series(self._h, ptrs, lens, n), "nma_set_series")

    def _idx(self, idx) -> torch.Tensor:
        if isinstance(idx, torch.Tensor):
            return idx.to(device=self.device, dtype=torch.int64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(idx, dtype=np.int64))).to(self.device)

    def gather(self, idx):
        cfg = self.cfg
        idx = self._idx(idx)
        p = idx.numel()
        tf = torch.empty(p, cfg.L0, cfg.Cf, dtype=torch.float32, device=self.device)
        mask = torch.empty(p, cfg.D, cfg.B + 1, dtype=torch.float32, device=self.device)
        shift = torch.empty_like(mask)
        _lib.check(self._lib.nma_gather(self._h, _ptr(idx), p, _ptr(tf), _ptr(mask), _ptr(shift), _stream()),
                   "nma_gather")
        return tf, mask, shift

    def alloc_outputs(self, p: int) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        d = self.device
        return {
            "terms": torch.empty(p, 4, dtype=torch.float32, device=d),
            "lf": torch.empty(p, cfg.L(cfg.F), dtype=torch.float32, device=d),
            "grad_params": torch.empty(self.n_params, dtype=torch.float32, device=d),
            "grad_theta": torch.empty(p, cfg.dtheta, dtype=torch.float32, device=d),
            "flags": torch.empty(p, dtype=torch.int32, device=d),
        }

    def elbo_fwd_bwd(self, params: torch.Tensor, eps: Optional[torch.Tensor], theta: torch.Tensor, idx: torch.Tensor,
                     objective: int = 0, path_target: float = 0.0, out: Optional[Dict[str, torch.Tensor]] = None):
        """eps=None: the library draws the base noise (Philox, `set_seed`)."""
        p = idx.numel()
        assert params.is_cuda and theta.is_cuda and idx.is_cuda
        assert params.dtype == torch.float32 and params.numel() == self.n_params and params.is_contiguous()
        assert eps is None or (eps.is_cuda and eps.shape == (p, self.cfg.L0) and eps.is_contiguous() and eps.dtype == torch.float32)
        assert theta.shape == (p, self.cfg.dtheta) and theta.is_contiguous() and theta.dtype == torch.float32
        assert idx.dtype == torch.int64
        if out is None:
            out = self.alloc_outputs(p)
        _lib.check(self._lib.nma_elbo_fwd_bwd(
            self._h, _ptr(params), _ptr(eps), _ptr(theta), _ptr(idx), p, int(objective), float(path_target),
            _ptr(out["terms"]), _ptr(out["lf"]), _ptr(out["grad_params"]), _ptr(out["grad_theta"]),
            _ptr(out["flags"]), _stream()), "nma_elbo_fwd_bwd")
        return out

    def forward_paths(self, params, eps, theta, idx):
        p = idx.numel()
        terms = torch.empty(p, 4, dtype=torch.float32, device=self.device)
        lf = torch.empty(p, self.cfg.L(self.cfg.F), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.nma_forward_paths(self._h, _ptr(params), _ptr(eps), _ptr(theta), _ptr(idx), p,
                                               _ptr(terms), _ptr(lf), _stream()), "nma_forward_paths")
        return terms, lf

    # ------------------------------------------------------------------
    # the whole iteration (nma_train_step): one sess.run([train_step, merged]) of the reference
    # ------------------------------------------------------------------
    def set_theta_flow(self, flow, priors) -> None:
        """Registers the theta posterior (viforssms_b200.theta_flow.ThetaFlow: masks, permutations, base distribution)
        and the diagonal Gaussian prior [(mean, scale), ...] of AR.py:178-182 with the handle."""
        masks = torch.cat([torch.from_numpy(m).reshape(-1) for m in flow.masks_np]).float().to(self.device)
        perms = np.stack(flow.perms).astype(np.int32) if flow.perms else np.zeros((0, flow.d), dtype=np.int32)
        self._tf_masks, self._tf_perms = masks, torch.from_numpy(perms).to(self.device)      # borrowed by the handle
        self._tf_flow = flow
        relu = 0 if flow.act is torch.nn.functional.elu else 1
        d = self.cfg.dtheta
        assert flow.d == d and len(priors) == d
        pm = (ctypes.c_float * d)(*[float(m) for m, _ in priors])
        ps = (ctypes.c_float * d)(*[float(s) for _, s in priors])
        _lib.check(self._lib.nma_set_theta_flow(self._h, _ptr(self._tf_masks), _ptr(self._tf_perms) if flow.nb > 1 else None,
                                                flow.nb, relu, flow.base_loc, flow.base_scale, pm, ps,
                                                int(getattr(flow, "softplus_out", False))), "nma_set_theta_flow")
        assert int(self._lib.nma_theta_flow_param_count(d, flow.nb)) == flow.n_params

    def set_fixed_theta(self, theta) -> None:
        """No theta posterior: theta is the constant `theta` in every row (the fixed-theta Lotka-Volterra script)."""
        d = self.cfg.dtheta
        assert len(theta) == d
        self._tf_flow = None
        pm = (ctypes.c_float * d)(*[float(t) for t in theta])
        ps = (ctypes.c_float * d)(*([1.0] * d))
        _lib.check(self._lib.nma_set_theta_flow(self._h, None, None, 0, 0, 0.0, 1.0, pm, ps, 0), "nma_set_theta_flow")

    def set_seed(self, seed: int, counter: int = 0) -> None:
        _lib.check(self._lib.nma_set_seed(self._h, int(seed), int(counter)), "nma_set_seed")

    def draw_counter(self) -> int:
        out = ctypes.c_uint64(0)
        _lib.check(self._lib.nma_get_counter(self._h, ctypes.byref(out)), "nma_get_counter")
        return int(out.value)

    def train_step(self, blob, grad, m, v, idx, scalars, *, objective=0, path_target=0.0, prior_on=True, obs_in_elbo=True,
                   tf_mask_grad=True, lr=1e-3, beta1=0.95, beta2=0.999, eps=1e-8, clip=0.0, theta_out=None,
                   lf_out=None) -> None:
        """nma_train_step on flat fp32 tensors (NMA variables ++ theta-posterior variables); everything stays on the
        device, `scalars` [8] receives the logged means (see nma_b200.h)."""
        n = self.n_params + (self._tf_flow.n_params if self._tf_flow is not None else 0)
        for t in (blob, grad, m, v):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == n
        assert idx.is_cuda and idx.dtype == torch.int64 and scalars.numel() >= 8 and scalars.dtype == torch.float32
        o = _lib.StepOpts(int(objective), float(path_target), int(bool(prior_on)), int(bool(obs_in_elbo)),
                          int(bool(tf_mask_grad)), float(lr), float(beta1), float(beta2), float(eps), float(clip))
        _lib.check(self._lib.nma_train_step(self._h, _ptr(blob), _ptr(grad), _ptr(m), _ptr(v), _ptr(idx), idx.numel(),
                                            ctypes.byref(o), _ptr(scalars), _ptr(theta_out), _ptr(lf_out), _stream()),
                   "nma_train_step")

    def step_buffers(self, p: int) -> Dict[str, torch.Tensor]:
        """Copies of what the last train_step left in the workspace: eps, z0, theta, logq_theta, terms, row_elbo."""
        ptrs = [c_void_p() for _ in range(6)]
        _lib.check(self._lib.nma_step_buffers(self._h, *[ctypes.byref(q) for q in ptrs]), "nma_step_buffers")
        cfg = self.cfg
        shapes = {"eps": (p, cfg.L0), "z0": (p, cfg.dtheta), "theta": (p, cfg.dtheta), "logq_theta": (p,), "terms": (p, 4),
                  "row_elbo": (p,)}
        out = {}
        for (name, shape), q in zip(shapes.items(), ptrs):
            out[name] = torch.as_tensor(_DevView(q.value, shape), device=self.device).clone()
        return out

    # ------------------------------------------------------------------
    # multi-GPU: the library's own NCCL communicator for the gradient all-reduce
    # ------------------------------------------------------------------
    def comm_create(self, rank: int, world: int, group=None) -> None:
        """Collective over `group` (any torch.distributed backend): rank 0 makes the NCCL id, everyone joins."""
        import torch.distributed as dist
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(self._lib.nma_comm_unique_id(buf), "nma_comm_unique_id")
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        backend = dist.get_backend(group)
        if backend == "nccl":
            t = t.to(self.device)
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        with torch.cuda.device(self.device):
            _lib.check(self._lib.nma_comm_create(self._h, raw, int(rank), int(world)), "nma_comm_create")

    def comm_destroy(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nma_comm_destroy(self._h)

    def comm_wait(self) -> None:
        _lib.check(self._lib.nma_comm_wait(self._h, _stream()), "nma_comm_wait")

    def comm_allreduce(self, t: torch.Tensor) -> None:
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        _lib.check(self._lib.nma_comm_allreduce(self._h, _ptr(t), t.numel(), _stream()), "nma_comm_allreduce")

    def adamax_step(self, params, grads, m, v, lr, beta1, beta2=0.999, eps=1e-8, clip=0.0) -> torch.Tensor:
        """In-place clip + Adamax on flat fp32 tensors; returns the 1-element global-norm tensor (device)."""
        n = params.numel()
        _lib.check(self._lib.nma_adamax_step(_ptr(params), _ptr(grads), _ptr(m), _ptr(v), n, float(lr), float(beta1),
                                             float(beta2), float(eps), float(clip), _ptr(self._norm),
                                             _ptr(self._scratch), _stream()), "nma_adamax_step")
        return self._norm


def tc_conv_raw(x: torch.Tensor, w: torch.Tensor, mode: int = 0, nacc: int = 2) -> torch.Tensor:
    """Test hook: bare tensor-core contraction.  x [Q,56] fp32, w [K,51,50] -> [Q,64] (see nma_b200.h)."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 56
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape[1:]) == (51, 50)
    out = torch.empty(x.shape[0], 64, dtype=torch.float32, device=x.device)
    _lib.check(lib.nma_tc_conv_raw(_ptr(x), _ptr(w), mode, nacc, _ptr(out), x.shape[0], w.shape[0], _stream()),
               "nma_tc_conv_raw")
    return out


def tc_wgrad_raw(x: torch.Tensor, da: torch.Tensor, K: int, bf16: bool = False) -> torch.Tensor:
    """Test hook: bare tensor-core weight gradient.  x, da [Q,56] fp32 -> [K,51,50] (see nma_b200.h);
    bf16=True: the 2-term bf16 split kernel."""
    lib = _lib.load()
    assert x.is_cuda and da.is_cuda and x.shape == da.shape and x.shape[1] == 56
    assert x.dtype == torch.float32 and da.dtype == torch.float32 and x.is_contiguous() and da.is_contiguous()
    gw = torch.zeros(K, 51, 50, dtype=torch.float32, device=x.device)
    fn = lib.nma_tc_wgrad_raw_bf if bf16 else lib.nma_tc_wgrad_raw
    _lib.check(fn(_ptr(x), _ptr(da), _ptr(gw), x.shape[0], K, _stream()), "nma_tc_wgrad_raw")
    return gw


def scan_ar1(z: torch.Tensor, x0: float, a: float, b: float, c: float) -> torch.Tensor:
    """x[0]=x0, x[i] = a*x[i-1] + b + c*z[i-1] on device (float64); AR_dat_gen.py:11-14."""
    lib = _lib.load()
    assert z.is_cuda and z.dtype == torch.float64 and z.is_contiguous()
    n = z.numel()
    x = torch.empty(n + 1, dtype=torch.float64, device=z.device)
    scratch = torch.empty((int(lib.nma_scan_scratch_bytes(n)) + 7) // 8, dtype=torch.float64, device=z.device)
    _lib.check(lib.nma_scan_ar1(_ptr(z), _ptr(x), n, x0, a, b, c, _ptr(scratch), scratch.numel() * 8, _stream()),
               "nma_scan_ar1")
    return x


def scan_affine(A: torch.Tensor, D: torch.Tensor, x0: float) -> torch.Tensor:
    """x[0]=x0, x[i] = A[i-1]*x[i-1] + D[i-1] on device (float64): the per-element form of the A12 scan."""
    lib = _lib.load()
    assert A.is_cuda and D.is_cuda and A.dtype == torch.float64 and D.dtype == torch.float64
    A, D = A.contiguous(), D.contiguous()
    n = A.numel()
    assert D.numel() == n
    x = torch.empty(n + 1, dtype=torch.float64, device=A.device)
    scratch = torch.empty((int(lib.nma_scan_scratch_bytes(n)) + 7) // 8, dtype=torch.float64, device=A.device)
    _lib.check(lib.nma_scan_affine(_ptr(A), _ptr(D), _ptr(x), n, float(x0), _ptr(scratch), scratch.numel() * 8, _stream()),
               "nma_scan_affine")
    return x


def time_till(obs: torch.Tensor, impute: int):
    """(obs_fill, obs_binary, time_till) on device from the n+1 noisy observations; AR_dat_gen.py:17-31."""
    lib = _lib.load()
    assert obs.is_cuda and obs.dtype == torch.float64 and obs.is_contiguous()
    n = obs.numel() - 1
    m = ((n - impute) // impute + 1) * impute
    fill = torch.empty(m, dtype=torch.float64, device=obs.device)
    binary = torch.empty_like(fill)
    till = torch.empty_like(fill)
    _lib.check(lib.nma_time_till(_ptr(obs), n, impute, _ptr(fill), _ptr(binary), _ptr(till), _stream()),
               "nma_time_till")
    return fill, binary, till


def rolling_var(x: torch.Tensor, K: int) -> torch.Tensor:
    """[np.var(x[i:i+K]) for i in range(len(x) - K)] on the device, float32, bit-exact with numpy (SV_dense.py:159-170)."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 1
    out = torch.empty(x.numel() - K, dtype=torch.float32, device=x.device)
    _lib.check(lib.nma_rolling_var(_ptr(x), x.numel(), int(K), _ptr(out), _stream()), "nma_rolling_var")
    return out


def philox_normal(n: int, seed: int, counter: int, stream_id: int, loc: float = 0.0, scale: float = 1.0,
                  device=None) -> torch.Tensor:
    """The normals the library draws for (seed, counter): stream 0 = eps, 1 = the theta posterior's base sample."""
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.float32, device=device if device is not None else "cuda")
    _lib.check(lib.nma_philox_normal(_ptr(out), n, int(seed), int(counter), int(stream_id), float(loc), float(scale),
                                     _stream()), "nma_philox_normal")
    return out


class DeviceThetaFlow:
    """The theta posterior on the device (nma_theta_flow_fwd / _bwd; A11, AR.py:376-391): two launches instead of the
    ~400 tiny ones of the host autograd module (which remains the test reference, tests/test_gpu_lvr_theta.py)."""

    def __init__(self, flow, device):
        self.flow = flow
        self.device = device
        self._lib = _lib.load()
        self.masks = torch.cat([torch.from_numpy(m).reshape(-1) for m in flow.masks_np]).float().to(device)
        perms = np.stack(flow.perms).astype(np.int32) if flow.perms else np.zeros((0, flow.d), dtype=np.int32)
        self.perms = torch.from_numpy(perms).to(device)
        self.relu = 0 if flow.act is torch.nn.functional.elu else 1

    def forward(self, params: torch.Tensor, z0: torch.Tensor):
        f = self.flow
        p = z0.shape[0]
        theta = torch.empty(p, f.d, dtype=torch.float32, device=self.device)
        logq = torch.empty(p, dtype=torch.float32, device=self.device)
        _lib.check(self._lib.nma_theta_flow_fwd(_ptr(params), _ptr(self.masks), _ptr(self.perms), _ptr(z0), p, f.d, f.nb,
                                                self.relu, f.base_loc, f.base_scale, _ptr(theta), _ptr(logq), _stream()),
                   "nma_theta_flow_fwd")
        return theta, logq

    def backward(self, params: torch.Tensor, z0: torch.Tensor, g_theta: torch.Tensor, g_logq: Optional[torch.Tensor],
                 g_params: torch.Tensor) -> None:
        """Accumulates d loss / d params into `g_params` (same length as `params`)."""
        f = self.flow
        p = z0.shape[0]
        _lib.check(self._lib.nma_theta_flow_bwd_ex(_ptr(params), _ptr(self.masks), _ptr(self.perms), _ptr(z0), p, f.d, f.nb,
                                                   self.relu, _ptr(g_theta), _ptr(g_logq) if g_logq is not None else None,
                                                   0.0, int(f.tf_mask_grad), _ptr(g_params), None, _stream()),
                   "nma_theta_flow_bwd_ex")
