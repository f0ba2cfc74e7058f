"""ctypes binding of libnma_b200.so (the C-ABI in include/nma_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present the
import / the first call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_uint32, c_void_p

from .config import CConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnma_b200.so")

EXPORTS = [
    "nma_last_error", "nma_version", "nma_create", "nma_destroy", "nma_param_count", "nma_param_layout",
    "nma_workspace_bytes", "nma_set_series", "nma_gather", "nma_elbo_fwd_bwd", "nma_forward_paths",
    "nma_adamax_step", "nma_scan_ar1", "nma_time_till", "nma_launch_stage", "nma_launch_count",
    "nma_set_tensor_cores", "nma_get_tensor_cores", "nma_tc_conv_raw", "nma_tc_wgrad_raw", "nma_tc_wgrad_raw_bf", "nma_rolling_var", "nma_theta_flow_fwd", "nma_theta_flow_bwd",
    "nma_theta_flow_bwd_ex", "nma_theta_flow_constrain", "nma_set_theta_flow", "nma_theta_flow_param_count", "nma_train_step",
    "nma_set_seed", "nma_get_counter", "nma_philox_normal", "nma_step_buffers",
    "nma_comm_unique_id", "nma_comm_create", "nma_comm_init", "nma_comm_destroy", "nma_comm_world", "nma_comm_wait",
    "nma_comm_allreduce", "nma_scan_scratch_bytes", "nma_scan_affine",
]


class StepOpts(ctypes.Structure):
    """nma_step_opts of include/nma_b200.h."""
    _fields_ = [("objective", c_int32), ("path_target", c_float), ("prior_on", c_int32), ("obs_in_elbo", c_int32),
                ("tf_mask_grad", c_int32), ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("clip", c_float)]

_lib = None


class NMAError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NMAError(
            f"{LIB_PATH} not found: build it with `python -m viforssms_b200.build` "
            "(there is no CPU fallback for the NMA ELBO step)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.nma_last_error.restype = c_char_p
    lib.nma_version.restype = c_int32
    lib.nma_create.argtypes = [POINTER(CConfig), POINTER(c_void_p)]
    lib.nma_create.restype = c_int32
    lib.nma_destroy.argtypes = [c_void_p]
    lib.nma_param_count.argtypes = [c_void_p]
    lib.nma_param_count.restype = c_int64
    lib.nma_workspace_bytes.argtypes = [c_void_p]
    lib.nma_workspace_bytes.restype = c_int64
    lib.nma_param_layout.argtypes = [c_void_p, POINTER(c_int64), c_int32]
    lib.nma_param_layout.restype = c_int32
    lib.nma_set_series.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_int64), c_int32]
    lib.nma_set_series.restype = c_int32
    lib.nma_gather.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nma_gather.restype = c_int32
    lib.nma_elbo_fwd_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nma_elbo_fwd_bwd.restype = c_int32
    lib.nma_forward_paths.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                      c_void_p]
    lib.nma_forward_paths.restype = c_int32
    lib.nma_adamax_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float,
                                    c_float, c_float, c_void_p, c_void_p, c_void_p]
    lib.nma_adamax_step.restype = c_int32
    lib.nma_launch_stage.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                     c_void_p]
    lib.nma_launch_stage.restype = c_int32
    lib.nma_launch_count.restype = c_int64
    lib.nma_set_tensor_cores.argtypes = [c_void_p, c_int32]
    lib.nma_set_tensor_cores.restype = c_int32
    lib.nma_get_tensor_cores.argtypes = [c_void_p]
    lib.nma_get_tensor_cores.restype = c_int32
    lib.nma_tc_conv_raw.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_int32, c_void_p]
    lib.nma_tc_conv_raw.restype = c_int32
    lib.nma_tc_wgrad_raw.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]
    lib.nma_tc_wgrad_raw.restype = c_int32
    lib.nma_tc_wgrad_raw_bf.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]
    lib.nma_tc_wgrad_raw_bf.restype = c_int32
    lib.nma_scan_ar1.argtypes = [c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double, c_void_p,
                                 c_int64, c_void_p]
    lib.nma_scan_ar1.restype = c_int32
    lib.nma_scan_scratch_bytes.argtypes = [c_int64]
    lib.nma_scan_scratch_bytes.restype = c_int64
    lib.nma_scan_affine.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_int64, c_void_p]
    lib.nma_scan_affine.restype = c_int32
    lib.nma_time_till.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nma_time_till.restype = c_int32
    lib.nma_rolling_var.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.nma_rolling_var.restype = c_int32
    lib.nma_theta_flow_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_float, c_float, c_void_p, c_void_p, c_void_p]
    lib.nma_theta_flow_fwd.restype = c_int32
    lib.nma_theta_flow_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nma_theta_flow_bwd.restype = c_int32
    lib.nma_theta_flow_bwd_ex.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                          c_void_p, c_void_p, c_float, c_int32, c_void_p, c_void_p, c_void_p]
    lib.nma_theta_flow_bwd_ex.restype = c_int32
    lib.nma_theta_flow_constrain.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    lib.nma_theta_flow_constrain.restype = c_int32
    lib.nma_set_theta_flow.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_float,
                                       POINTER(c_float), POINTER(c_float), c_int32]
    lib.nma_set_theta_flow.restype = c_int32
    lib.nma_theta_flow_param_count.argtypes = [c_int32, c_int32]
    lib.nma_theta_flow_param_count.restype = c_int64
    lib.nma_train_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, POINTER(StepOpts),
                                   c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nma_train_step.restype = c_int32
    lib.nma_set_seed.argtypes = [c_void_p, ctypes.c_uint64, ctypes.c_uint64]
    lib.nma_set_seed.restype = c_int32
    lib.nma_get_counter.argtypes = [c_void_p, POINTER(ctypes.c_uint64)]
    lib.nma_get_counter.restype = c_int32
    lib.nma_philox_normal.argtypes = [c_void_p, c_int64, ctypes.c_uint64, ctypes.c_uint64, c_uint32, c_float, c_float,
                                      c_void_p]
    lib.nma_philox_normal.restype = c_int32
    lib.nma_step_buffers.argtypes = [c_void_p] + [POINTER(c_void_p)] * 6
    lib.nma_step_buffers.restype = c_int32
    lib.nma_comm_unique_id.argtypes = [ctypes.c_char_p]
    lib.nma_comm_unique_id.restype = c_int32
    lib.nma_comm_create.argtypes = [c_void_p, ctypes.c_char_p, c_int32, c_int32]
    lib.nma_comm_create.restype = c_int32
    lib.nma_comm_init.argtypes = [c_void_p, c_void_p]
    lib.nma_comm_init.restype = c_int32
    lib.nma_comm_destroy.argtypes = [c_void_p]
    lib.nma_comm_destroy.restype = c_int32
    lib.nma_comm_world.argtypes = [c_void_p]
    lib.nma_comm_world.restype = c_int32
    lib.nma_comm_wait.argtypes = [c_void_p, c_void_p]
    lib.nma_comm_wait.restype = c_int32
    lib.nma_comm_allreduce.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.nma_comm_allreduce.restype = c_int32
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().nma_last_error()
        raise NMAError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
