"""ctypes binding of ``libhp_b200.so`` (C ABI: ``include/hp_b200.h``).

There is no CPU or eager-PyTorch fallback behind these calls: if the library is missing, or a
tensor is not a CUDA tensor, the call raises.  PyTorch only supplies device memory, the
current stream and (in ``dist.py``) the process group.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhp_b200.so")

# enums of include/hp_b200.h
LOSS_MSE, LOSS_KL = 1, 2
PIPE_OVERLAP_PREV = 1          # HP_PIPE_OVERLAP_PREV (include/hp_b200.h)
PIPE_DEFER_EXCHANGE = 2        # HP_PIPE_DEFER_EXCHANGE


def pipe_flags(overlap, defer=False):
    """False/0 -> serialised launch; True -> overlapped at the library's default depth; int d in 1..8 ->
    HP_PIPE_OVERLAP_PREV | HP_PIPE_DEPTH(d).  ``defer`` adds HP_PIPE_DEFER_EXCHANGE (sharded steps)."""
    extra = PIPE_DEFER_EXCHANGE if defer else 0
    if overlap is True:
        return PIPE_OVERLAP_PREV | extra
    d = int(overlap or 0)
    if d == 0:
        return extra
    if not 1 <= d <= 8:
        raise ValueError(f"overlap depth must be 1..8, got {overlap!r}")
    return PIPE_OVERLAP_PREV | (d << 8) | extra
PLG_BASE, PLG_ONE_MINUS = 0, 1
RD_BASE, RD_X1, RD_X5, RD_X6, RD_RD4 = 0, 1, 2, 3, 4
MODE_MIN, MODE_MAX = 0, 1
GRAD_SCALAR, GRAD_PER_MAP, GRAD_PER_SAMPLE = 0, 1, 2
MAX_K = 64
PEER_MAX_K = 27                 # sharded with collective='peer': 2*(4+2K+6) <= 128 mailbox entries per source

_vp, _i, _f, _d, _sz = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t

#: name -> (restype, argtypes) for every symbol declared in include/hp_b200.h
PROTOTYPES = {
    "hp_version": (_i, []),
    "hp_last_error": (C.c_char_p, []),
    "hp_device_sm_count": (_i, []),
    "hp_workspace_bytes": (_sz, [_i, _i]),
    "hp_argmax_decode": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hp_soft_argmax": (_i, [_vp, _i, _i, _i, _f, _f, _vp, _vp]),
    "hp_pck_accumulate": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp]),
    "hp_pck_finalize": (_i, [_vp, _i, _vp, _vp]),
    "hp_accuracy": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp]),
    "hp_gaussian_target": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp]),
    "hp_mse_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hp_mse_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "hp_kl_fwd": (_i, [_vp, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_kl_bwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "hp_pseudo_label": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "hp_regdisp_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _i, _i, _i, _i, _i, _i, _vp,
                            _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_regdisp_bwd": (_i, [_vp, _vp, _vp, _i, _i, _f, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "hp_regdisp_fwd_heads": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _i, _i, _f, _vp, _i, _i, _f, _i, _i, _i, _i, _i, _i, _i,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_regdisp_bwd_heads": (_i, [_vp, _vp, _i, _i, _f, _vp, _i, _i, _f, _vp, _i, _i, _f, _i, _i, _i, _i, _i, _vp, _vp, _vp,
                                  _vp, _i, _vp, _vp]),
    "hp_regdisp_materialize": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "hp_fuse_multiscale": (_i, [_vp, _i, _i, _f, _vp, _i, _i, _f, _vp, _f, _i, _i, _i, _vp, _vp]),
    "hp_fuse_multiscale_pair": (_i, [_vp, _i, _i, _f, _vp, _i, _i, _f, _i, _i, _i, _vp, _f, _i, _i, _vp, _vp]),
    "hp_fuse_decode_pck": (_i, [_vp, _i, _i, _f, _vp, _i, _i, _f, _vp, _f, _vp, _i, _i, _i, _i, _d,
                                _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_pipeline_fused": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _i, _vp, _f, _d, _i,
                               _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "hp_pipeline_fused_ex": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _i, _vp, _f, _d, _i,
                                  _vp, _vp, _vp, _vp, _i, _vp, _vp, C.c_uint, _vp]),
    "hp_pipeline_fused_peer": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _i, _vp, _f, _d, _i,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.c_int64, C.c_uint, _vp]),
    "hp_pipeline_plan_create": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _i, _vp, _f, _d, _i,
                                     _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, C.c_uint, _vp]),
    "hp_pipeline_plan_launch": (_i, [_vp, _vp]),
    "hp_pipeline_plan_destroy": (_i, [_vp]),
    "hp_debug_pipeline_trace_words": (_sz, []),
    "hp_debug_pipeline_trace": (_i, [_vp, _sz]),
    "hp_debug_regdisp_trace_words": (_sz, []),
    "hp_debug_regdisp_trace": (_i, [_vp, _sz]),
    "hp_mse0_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hp_mse0_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "hp_kl5_fwd": (_i, [_vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_kl5_bwd": (_i, [_vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "hp_label_fusion": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "hp_argmax_decode_f64": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "hp_refine_quarter": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "hp_group_accuracy": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "hp_pipeline_finalize": (_i, [_vp, _i, _vp, _vp]),
    "hp_peer_mailbox_bytes": (_sz, [_i]),
    "hp_peer_alloc": (_i, [_i, _vp]),
    "hp_peer_free": (_i, [_vp]),
    "hp_peer_export": (_i, [_vp, _vp]),
    "hp_peer_import": (_i, [_vp, _vp]),
    "hp_peer_close": (_i, [_vp]),
    "hp_pipeline_finalize_peer": (_i, [_vp, _vp, _i, _i, _i, C.c_int64, _vp, _vp, _vp]),
    "hp_pipeline_flush_peer": (_i, [_vp, _vp, _i, _i, _vp]),
    "hp_pck_finalize_peer": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "hp_fuse_decode_pck_peer": (_i, [_vp, _i, _i, _f, _vp, _i, _i, _f, _vp, _f, _vp, _i, _i, _i, _i, _d,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.c_uint, _vp, _vp, _vp]),
    "hp_pipeline_fused_host": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _i, _vp, _f, _d, _i, _i,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once).  Raises loudly when it is absent - build it with
    ``python __graft_entry__.py`` / ``__graft_entry__.build()``."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the B200 heatmap path has no CPU fallback. "
                    "Build it with `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)          # AttributeError if the .so is stale
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


_fn_cache = {}


def call(name, *args):
    """Invoke an ``int``-returning entry point and raise ``RuntimeError(hp_last_error())`` on failure.
    (Bound functions are cached: attribute lookup on a CDLL costs about as much as a small kernel launch.)"""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {load().hp_last_error().decode(errors='replace')}")


# ------------------------------------------------------------------------------------------
# tensor plumbing
# ------------------------------------------------------------------------------------------

def ptr(t):
    """Device address as a plain int (ctypes converts it through the prototype's c_void_p), None -> NULL."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    """Raw handle of PyTorch's current stream on ``device`` (one C call; ``torch.cuda.current_stream`` builds a
    Stream object and costs several microseconds per launch)."""
    idx = device.index if device is not None and device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class on_device:
    """``with on_device(dev):`` - make ``dev`` current for the launches inside; free when it already is
    (the common one-process-per-GPU case), unlike ``torch.cuda.device`` which always round-trips the driver."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index
        self.prev = -1

    def __enter__(self):
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
            self.prev = -1
        return False


def require_cuda(t: torch.Tensor, what: str, dtype=torch.float32) -> torch.Tensor:
    """Contiguous CUDA tensor of ``dtype`` or an exception - never a silent host path."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor is on {t.device}; the B200 heatmap path runs on CUDA only "
                           "(there is no CPU fallback)")
    if t.dtype != dtype:
        if dtype == torch.float32 and t.dtype in (torch.float16, torch.bfloat16):
            t = t.float()                     # exact widening
        else:
            raise TypeError(f"{what}: expected {dtype}, got {t.dtype}")
    return t.contiguous()


_ws_cache = {}
_ws_need = None


def workspace(device: torch.device, n_maps: int, K: int) -> torch.Tensor:
    """Zero-initialised scratch (block counter, PCK counters, per-map partials) per (device, stream).
    Kernels restore the zero state before they exit, so it is reused without a memset."""
    global _ws_need
    if _ws_need is None:
        _ws_need = int(load().hp_workspace_bytes(int(n_maps), int(K)))   # a constant of the library (header + slack)
    need = _ws_need
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch._C._cuda_getCurrentRawStream(idx))
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


# ------------------------------------------------------------------------------------------
# Gaussian table: tab[d2] = exp(-d2 / (2 sigma^2)), built with the reference's own expression
# ------------------------------------------------------------------------------------------
_tab_cache = {}


def integer_tmp(tmp_size) -> int:
    if float(tmp_size) != int(tmp_size) or tmp_size < 0:
        raise NotImplementedError(f"patch half-width {tmp_size} is not a non-negative integer "
                                  "(sigma*3, sigma*2 or sigma*1.5 must be integral)")
    return int(tmp_size)


def gaussian_table_host(sigma, tmp: int) -> np.ndarray:
    """uda/dataset/util.py:49-54 / regda_4.py:56-61 evaluated with numpy in float32, then indexed
    by d2 = dx^2 + dy^2 (the value depends on d2 only).  Same host + same numpy => the generated
    maps are bit-equal to the reference's."""
    size = 2 * tmp + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    d2f = (x - x0) ** 2 + (y - y0) ** 2
    g = np.exp(-d2f / (2 * sigma ** 2))
    d2 = d2f.astype(np.int64)
    tab = np.zeros(2 * tmp * tmp + 1, dtype=np.float32)
    tab[d2.ravel()] = g.ravel().astype(np.float32)
    # every pixel with the same d2 must carry the same bits (numpy's SIMD exp is lane independent);
    # holes (d2 values that are not a sum of two squares) are never indexed by the kernels
    if not np.array_equal(tab[d2], g.astype(np.float32)):
        raise RuntimeError("numpy exp is not a pure function of its argument on this host")
    return tab


def gaussian_table(sigma, tmp: int, device: torch.device) -> torch.Tensor:
    key = (float(sigma), int(tmp), str(device))
    t = _tab_cache.get(key)
    if t is None:
        t = torch.from_numpy(gaussian_table_host(sigma, tmp)).to(device)
        _tab_cache[key] = t
    return t
