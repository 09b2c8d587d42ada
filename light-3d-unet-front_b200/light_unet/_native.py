"""ctypes binding of libl3d.so (the C-ABI declared in include/l3d.h).

There is NO fallback: if the library is missing or a CUDA device is not used,
the product path raises.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_double, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("L3D_LIB", os.path.join(_PKG_DIR, "libl3d.so"))

L3D_F32, L3D_F16 = 0, 1
ABI_VERSION = 5


class Act(Structure):
    _fields_ = [("ptr", c_void_p), ("C", c_int32), ("ldc", c_int32), ("dtype", c_int32), ("pad_", c_int32)]


class Norm(Structure):
    _fields_ = [("stats", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("drop", c_void_p),
                ("eps", c_float), ("slope", c_float), ("count", c_int32), ("pad_", c_int32)]


class NativeError(RuntimeError):
    pass


_lib = None

# name -> argtypes ; every function returns int status except the ones listed in _SPECIAL
_P = c_void_p
_SIGS = {
    "l3d_dwpw_fwd": [POINTER(Act), POINTER(Norm), c_int, c_int, c_int, c_int, _P, _P, _P,
                     POINTER(Act), _P, POINTER(Act), _P, POINTER(Act), _P],
    "l3d_merge_fwd_rank1": [POINTER(Act), POINTER(Norm), POINTER(Act), _P, POINTER(Norm), c_int, c_int, c_int, c_int, c_float,
                            POINTER(Act), POINTER(Act), _P],
    "l3d_dw_c1_fwd": [POINTER(Act), POINTER(Norm), c_int, c_int, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P, _P],
    "l3d_dwpw_fwd_rank1": [_P, _P, c_int, POINTER(Norm), c_int, c_int, c_int, c_int, _P, _P, POINTER(Act), _P, _P],
    "l3d_dwpw_fwd2": [POINTER(Act), POINTER(Act), POINTER(Norm), c_int, c_int, c_int, c_int, _P, _P, _P,
                      POINTER(Act), _P, POINTER(Act), _P, _P],
    "l3d_conv3_fwd": [POINTER(Act), POINTER(Norm), c_int, c_int, c_int, c_int, _P, c_int, POINTER(Act), _P, _P, POINTER(Act), _P, _P],
    "l3d_merge_fwd": [POINTER(Act), POINTER(Norm), POINTER(Act), POINTER(Norm), c_int, c_int, c_int, c_int, c_float,
                      POINTER(Act), POINTER(Act), _P, _P, c_int, _P, _P, _P],
    "l3d_convt_fwd": [POINTER(Act), c_int, c_int, c_int, c_int, _P, _P, POINTER(Act), c_int, c_int, c_int,
                      c_int, c_int, c_int, _P],
    "l3d_merge_bwd": [POINTER(Act), POINTER(Act), POINTER(Act), POINTER(Act), POINTER(Act), POINTER(Norm),
                      POINTER(Act), POINTER(Norm), c_int, c_int, c_int, c_int, c_float,
                      _P, c_int, _P, _P, _P, _P, POINTER(Act), _P, _P, _P],
    "l3d_pw_bwd": [POINTER(Act), POINTER(Act), POINTER(Norm), _P, POINTER(Act), POINTER(Norm),
                   c_int, c_int, c_int, c_int, _P, _P, POINTER(Act), c_int, _P],
    "l3d_dw_bwd": [POINTER(Act), POINTER(Act), POINTER(Norm), c_int, c_int, c_int, c_int, _P, _P,
                   POINTER(Act), c_int, _P, _P],
    "l3d_conv3_bwd": [POINTER(Act), POINTER(Act), POINTER(Norm), _P, POINTER(Act), POINTER(Norm),
                      c_int, c_int, c_int, c_int, _P, c_int, _P, POINTER(Act), c_int, _P, _P, c_int64, _P],
    "l3d_convt_bwd": [POINTER(Act), c_int, c_int, c_int, c_int, c_int, c_int, POINTER(Act), c_int, c_int, c_int, c_int,
                      _P, _P, _P, POINTER(Act), c_int, _P],
    "l3d_norm_param_grad": [_P, c_int, c_int, _P, _P, _P],
    "l3d_norm_param_grad_batch": [c_int, _P, _P, _P, _P, c_int, _P],
    "l3d_ftl_sums": [_P, _P, c_int64, _P, _P],
    "l3d_ftl_finish": [_P, c_float, c_float, c_float, c_float, _P, _P, _P],
    "l3d_ftl_grad": [_P, c_int64, _P, _P, _P, _P],
    "l3d_gather_windows": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P],
    "l3d_stitch": [_P, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, _P, _P,
                   c_float, _P, _P],
    "l3d_stitch_slab": [_P, _P, c_int, _P, c_int, _P, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P,
                        c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, _P],
    "l3d_threshold": [_P, c_int64, c_float, _P, _P],
    "l3d_ccl_label": [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "l3d_bbox_init": [_P, c_int, _P],
    "l3d_bbox_reduce": [_P, _P, c_int, c_int, c_int, _P, c_int, _P],
    "l3d_label_pair_stats": [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "l3d_patch_extract": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P],
    "l3d_patch_flip": [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "l3d_patch_rotate": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "l3d_patch_zoom": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    "l3d_patch_intensity": [_P, _P, _P, _P, _P, c_int, c_int64, _P],
    "l3d_tc_selftest": [_P, _P, c_int, c_int, c_int, _P, _P],
    "l3d_conv3_debug_read": [_P, c_int],
    "l3d_tc_selftest_tf32": [_P, _P, c_int, c_int, c_int, c_int, _P, _P],
    "l3d_tc_selftest_mn16": [_P, _P, c_int, c_int, c_int, _P, _P],
}
EXPORTS = sorted(list(_SIGS) + ["l3d_add_launch_count", "l3d_env_refresh", "l3d_last_kernel", "l3d_last_error", "l3d_abi_version", "l3d_launch_count", "l3d_ccl_workspace_elems",
                                "l3d_conv3_bwd_workspace_bytes"])


def lib():
    """Load libl3d.so (once).  Raises NativeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (or "
            f"`python light-3d-unet-front_b200/build.py`). There is no CPU / PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.l3d_last_error.restype = ctypes.c_char_p
    L.l3d_last_error.argtypes = []
    L.l3d_last_kernel.restype = ctypes.c_char_p
    L.l3d_last_kernel.argtypes = []
    L.l3d_abi_version.restype = c_int
    L.l3d_add_launch_count.restype = None
    L.l3d_add_launch_count.argtypes = [c_int64]
    L.l3d_env_refresh.restype = None
    L.l3d_env_refresh.argtypes = []
    L.l3d_launch_count.restype = c_int64
    L.l3d_ccl_workspace_elems.restype = c_int64
    L.l3d_ccl_workspace_elems.argtypes = [c_int64]
    L.l3d_conv3_bwd_workspace_bytes.restype = c_int64
    L.l3d_conv3_bwd_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, c_int]
    for name, sig in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = c_int
        fn.argtypes = sig
    if L.l3d_abi_version() != ABI_VERSION:
        raise NativeError(f"libl3d ABI {L.l3d_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = L
    return L


class KernelTimer:
    """Optional per-entry-point device timing (CUDA events on the launching stream).  Used by bench.py to
    measure the dominant kernel's average launch duration inside a profiled pass; off by default."""

    def __init__(self):
        self.enabled = False
        self.records = []   # (name, tag, start_event, stop_event, algorithmic bytes, kernel)
        self.by_kernel = {}
        self.tag = ""

    def start(self):
        self.records.clear()
        self.enabled = True

    def stop(self):
        """-> {(name, tag): (launches, total_ms, total algorithmic bytes)}"""
        self.enabled = False
        torch.cuda.synchronize()
        out = {}
        self.by_kernel = {}     # kernel actually launched by a dispatching entry point -> (launches, ms, algorithmic bytes)
        for name, tag, e0, e1, nbytes, kern in self.records:
            dt = e0.elapsed_time(e1)
            n, ms, b = out.get((name, tag), (0, 0.0, 0))
            out[(name, tag)] = (n + 1, ms + dt, b + nbytes)
            if kern:
                n, ms, b = self.by_kernel.get(kern, (0, 0.0, 0))
                self.by_kernel[kern] = (n + 1, ms + dt, b + nbytes)
        self.records.clear()
        return out


_DISPATCHING = {"l3d_dwpw_fwd", "l3d_conv3_fwd", "l3d_convt_fwd", "l3d_dwpw_fwd_rank1", "l3d_dw_c1_fwd", "l3d_dwpw_fwd2"}
TIMER = KernelTimer()


def call(name: str, *args, algo_bytes: int = 0):
    """Invoke an entry point; non-zero status -> NativeError(l3d_last_error()).  `algo_bytes` (the launch's
    algorithmic HBM bytes: every operand read once, every result written once) is only recorded by TIMER."""
    L = lib()
    if TIMER.enabled:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L, name)(*args)
        e1.record()
        kern = L.l3d_last_kernel().decode() if name in _DISPATCHING else ""
        TIMER.records.append((name, TIMER.tag, e0, e1, int(algo_bytes), kern))
    else:
        rc = getattr(L, name)(*args)
    if rc != 0:
        raise NativeError(f"{name} failed ({rc}): {L.l3d_last_error().decode(errors='replace')}")


def refresh_env() -> None:
    """libl3d caches its L3D_* environment knobs per call site; call this after changing one in a running process."""
    lib().l3d_env_refresh()


def add_launch_count(n: int) -> None:
    """Launches replayed from a CUDA graph (they do not pass through the entry points that count)."""
    lib().l3d_add_launch_count(int(n))


def launch_count() -> int:
    return int(lib().l3d_launch_count())


def stream_ptr(device=None) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise NativeError(
            f"{what}: tensor is on {t.device}; the B200-native path runs on CUDA only (no CPU fallback)")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return L3D_F32
    if dt == torch.float16:
        return L3D_F16
    raise NativeError(f"unsupported activation dtype {dt}: the 16-bit storage format of libl3d is IEEE fp16 (torch.float16)")


def ptr(t) -> c_void_p:
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


_NULL_ACT = Act(None, 0, 0, 0, 0)


def act(t: torch.Tensor | None, ch_off: int = 0, C: int | None = None) -> Act:
    """View of a channels-last [N, D, H, W, Ctot] tensor restricted to channels [ch_off, ch_off+C)."""
    if t is None:
        return Act(None, 0, 0, 0, 0)
    assert t.is_contiguous(), "activation buffers must be contiguous NDHWC"
    ctot = t.shape[-1]
    if C is None:
        C = ctot - ch_off
    return Act(c_void_p(t.data_ptr() + ch_off * t.element_size()), C, ctot, dtype_code(t.dtype), 0)


def norm(stats=None, gamma=None, beta=None, drop=None, eps=1e-5, slope=1.0, count=1) -> Norm:
    if stats is None:
        return Norm(None, None, None, None, 0.0, 1.0, 1, 0)
    return Norm(c_void_p(stats.data_ptr()), c_void_p(gamma.data_ptr()), c_void_p(beta.data_ptr()),
                c_void_p(drop.data_ptr()) if drop is not None else None, eps, slope, int(count), 0)
