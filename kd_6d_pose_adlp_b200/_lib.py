"""ctypes binding of ``libkdot.so`` (C ABI declared in ``include/kdot.h``).

There is no CPU fallback: if the library is missing or no CUDA device is visible, every entry point
raises.  Build it with ``python -m kd_6d_pose_adlp_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

KDOT_LAYOUT_CELL_MAJOR = 0
KDOT_LAYOUT_SLOT_MAJOR = 1
KDOT_IMG_SKIPPED, KDOT_IMG_OK, KDOT_IMG_DEGENERATE, KDOT_IMG_TOO_MANY_ROUNDS = 0, 1, -1, -2
KDOT_MAX_ROUNDS = 1024

# every symbol include/kdot.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = (
    "kdot_sinkhorn_fwd_bwd", "kdot_kernel_mmd_fwd_bwd", "kdot_workspace_bytes", "kdot_workspace_bytes_ex", "kdot_host_ctx_create", "kdot_host_ctx_destroy",
    "kdot_sinkhorn_fwd_bwd_host", "kdot_host_ctx_last_traffic", "kdot_host_ctx_last_timing", "kdot_select_cells", "kdot_gather_decode_fwd", "kdot_gather_decode_bwd", "kdot_last_error",
    "kdot_version", "kdot_launch_count", "kdot_measure_fp32_peak_tflops", "kdot_debug_set_clock_buffer",
    "kdot_focal_loss_fwd_bwd", "kdot_focal_workspace_bytes", "kdot_reg3d_loss_fwd_bwd",
    "kdot_ssc_count", "kdot_ssc_pick", "kdot_ssc_assign",
)

_lib = None


class KdotError(RuntimeError):
    pass


def load(path: str):
    """dlopen a build of the C ABI and declare the prototypes of include/kdot.h on it."""
    if not os.path.exists(path):
        raise KdotError(
            f"{path} not found: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -m kd_6d_pose_adlp_b200.build`."
        )
    L = C.CDLL(path)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.kdot_last_error.restype = C.c_char_p
    L.kdot_version.restype = i32
    L.kdot_launch_count.restype = C.c_ulonglong
    L.kdot_workspace_bytes.restype = sz
    L.kdot_workspace_bytes.argtypes = [i32] * 5
    L.kdot_workspace_bytes_ex.restype = sz
    L.kdot_workspace_bytes_ex.argtypes = [i32] * 5 + [f32]
    L.kdot_sinkhorn_fwd_bwd.restype = i32
    L.kdot_sinkhorn_fwd_bwd.argtypes = (
        [vp] * 6 + [i32] * 6 + [f32] * 6 + [i32] + [vp] * 6 + [vp, sz, vp]
    )
    L.kdot_kernel_mmd_fwd_bwd.restype = i32
    L.kdot_kernel_mmd_fwd_bwd.argtypes = [vp] * 6 + [i32] * 7 + [f32] * 3 + [i32] + [vp] * 5 + [vp]
    L.kdot_host_ctx_create.restype = vp
    L.kdot_host_ctx_create.argtypes = [i32] * 6
    L.kdot_host_ctx_destroy.argtypes = [vp]
    L.kdot_host_ctx_last_traffic.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.kdot_host_ctx_last_timing.argtypes = [vp, C.POINTER(C.c_double)]
    L.kdot_sinkhorn_fwd_bwd_host.restype = i32
    L.kdot_sinkhorn_fwd_bwd_host.argtypes = [vp] * 7 + [i32] + [f32] * 6 + [i32, i32] + [vp] * 5
    L.kdot_select_cells.restype = i32
    L.kdot_select_cells.argtypes = (
        [vp] * 5 + [i32, vp, i32, i32, i32, f32, i32, f32, i32] + [vp] * 8 + [vp]
    )
    L.kdot_gather_decode_fwd.restype = i32
    L.kdot_gather_decode_fwd.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]
    L.kdot_gather_decode_bwd.restype = i32
    L.kdot_gather_decode_bwd.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]
    L.kdot_debug_set_clock_buffer.argtypes = [vp]
    L.kdot_focal_loss_fwd_bwd.restype = i32
    L.kdot_focal_loss_fwd_bwd.argtypes = [vp, vp, i32, i32, i32, vp, f32, f32, vp, vp, vp, sz, vp]
    L.kdot_focal_workspace_bytes.restype = sz
    L.kdot_focal_workspace_bytes.argtypes = []
    L.kdot_reg3d_loss_fwd_bwd.restype = i32
    L.kdot_reg3d_loss_fwd_bwd.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
    L.kdot_ssc_count.restype = i32
    L.kdot_ssc_count.argtypes = [vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, f32, vp, vp, vp, vp, vp]
    L.kdot_ssc_pick.restype = i32
    L.kdot_ssc_pick.argtypes = [vp, vp, i32, i32, i32, i32, C.c_uint64, vp, vp]
    L.kdot_ssc_assign.restype = i32
    L.kdot_ssc_assign.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.kdot_measure_fp32_peak_tflops.restype = C.c_double
    L.kdot_measure_fp32_peak_tflops.argtypes = [i32, i32]
    return L


def lib():
    """The loaded shared library (loads on first use; raises if it was not built)."""
    global _lib
    if _lib is None:
        _lib = load(os.environ.get("KDOT_LIB", LIB_PATH))  # KDOT_LIB: tuning aid, alternative builds of the same ABI
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise KdotError(f"{what} failed (code {rc}): {lib().kdot_last_error().decode()}")


def launch_count() -> int:
    return int(lib().kdot_launch_count())
