"""Multiscale heatmap fusion on the GPU.

The reference does this inline (``train1.py:410-424`` == ``test.py:362-376``) with three
``nn.Upsample(mode='bilinear')`` modules and two elementwise ops; :func:`fuse_multiscale` produces
the same two tensors with one gather+blend kernel each and no intermediate tensors."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _fuse(lo, a_lo, mid, a_mid, hi, a_hi, size):
    lo = _lib.require_cuda(lo.detach(), "fuse(lo)")
    B, K, hl, wl = lo.shape
    H, W = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
    hm = wm = 0
    if mid is not None:
        mid = _lib.require_cuda(mid.detach(), "fuse(mid)")
        if mid.shape[:2] != lo.shape[:2]:
            raise ValueError("fuse: batch/joint dims differ")
        hm, wm = mid.shape[2], mid.shape[3]
    if hi is not None:
        hi = _lib.require_cuda(hi.detach(), "fuse(hi)")
        if tuple(hi.shape) != (B, K, H, W):
            raise ValueError(f"fuse: hi is {tuple(hi.shape)}, expected {(B, K, H, W)}")
    out = torch.empty((B, K, H, W), dtype=torch.float32, device=lo.device)
    with _lib.on_device(lo.device):
        _lib.call("hp_fuse_multiscale", _lib.ptr(lo), hl, wl, C.c_float(a_lo), _lib.ptr(mid), hm, wm, C.c_float(a_mid),
                  _lib.ptr(hi), C.c_float(a_hi), B * K, H, W, _lib.ptr(out), _lib.stream_ptr(lo.device))
    return out


def upsample_bilinear(x, size):
    """``nn.Upsample(size=size, mode='bilinear')(x)`` (align_corners=False)."""
    return _fuse(x, 1.0, None, 0.0, None, 0.0, size)


def fuse_multiscale(y_adv3, y_adv2, size_hi=64, size_mid=32, lazy=False):
    """train1.py:410-424 -> ``(target5, target0)`` with
    ``target5 = 0.5*up_hi(y_adv3) + up_hi(y_adv2)`` and ``target0 = up_mid(y_adv3)`` (inputs detached).
    ``lazy=True``: neither map is built - ``(FusedHeads, LazyUpsample)``, which the disparity modules accept as ``y_adv2``
    (x6 'max' interpolates ``target5`` inside its loss kernel; anything else materialises on first use)."""
    if lazy:
        return FusedHeads(y_adv3, y_adv2, 0.5, 1.0, size_hi), LazyUpsample(y_adv3, size_mid)
    if isinstance(size_hi, int) and isinstance(size_mid, int) and 2 * size_mid == size_hi:
        # both maps from ONE launch (hp_fuse_multiscale_pair; any geometry it does not cover falls back to two inside the call)
        lo = _lib.require_cuda(y_adv3.detach(), "fuse(lo)")
        mid = _lib.require_cuda(y_adv2.detach(), "fuse(mid)")
        if mid.shape[:2] != lo.shape[:2]:
            raise ValueError("fuse: batch/joint dims differ")
        B, K, hl, wl = lo.shape
        target5 = torch.empty((B, K, size_hi, size_hi), dtype=torch.float32, device=lo.device)
        target0 = torch.empty((B, K, size_mid, size_mid), dtype=torch.float32, device=lo.device)
        with _lib.on_device(lo.device):
            _lib.call("hp_fuse_multiscale_pair", _lib.ptr(lo), hl, wl, C.c_float(0.5), _lib.ptr(mid), mid.shape[2], mid.shape[3],
                      C.c_float(1.0), B * K, size_hi, size_hi, _lib.ptr(target5), C.c_float(1.0), size_mid, size_mid,
                      _lib.ptr(target0), _lib.stream_ptr(lo.device))
        return target5, target0
    target5 = _fuse(y_adv3, 0.5, y_adv2, 1.0, None, 0.0, size_hi)
    target0 = _fuse(y_adv3, 1.0, None, 0.0, None, 0.0, size_mid)
    return target5, target0


def fuse_three_scales(lo, mid, hi):
    """BASELINE.json configs[3]: the same rule over three resolutions (e.g. 32/64/128):
    ``0.5*up(lo) + up(mid) + hi`` at the resolution of ``hi``."""
    return _fuse(lo, 0.5, mid, 1.0, hi, 1.0, (hi.shape[2], hi.shape[3]))



# ---- maps that are NOT built until somebody needs them --------------------------------------------------------------------
def _unlazy(x):
    if isinstance(x, _LazyMap):
        return x.materialise()
    if isinstance(x, (list, tuple)):
        return type(x)(_unlazy(v) for v in x)
    if isinstance(x, dict):
        return {k: _unlazy(v) for k, v in x.items()}
    return x


class _LazyMap:
    """Stands for a CUDA float32 tensor that has not been computed.  The consumers that can do better than reading it from
    memory (``RegressionDisparityx6`` for :class:`FusedHeads`) recognise the type; every other use - a torch function or
    method, arithmetic, any attribute of a tensor - sees :meth:`materialise` (one ``hp_fuse_multiscale`` launch, cached)."""

    _map = None

    def materialise(self):
        raise NotImplementedError

    def __getattr__(self, name):                 # only reached for names the lazy object itself does not have
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialise(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return func(*_unlazy(args), **_unlazy(kwargs or {}))

    def __add__(self, other):
        return self.materialise() + _unlazy(other)

    def __radd__(self, other):
        return _unlazy(other) + self.materialise()

    def __sub__(self, other):
        return self.materialise() - _unlazy(other)

    def __rsub__(self, other):
        return _unlazy(other) - self.materialise()

    def __mul__(self, other):
        return self.materialise() * _unlazy(other)

    def __rmul__(self, other):
        return _unlazy(other) * self.materialise()

    def __truediv__(self, other):
        return self.materialise() / _unlazy(other)

    def __neg__(self):
        return -self.materialise()

    def __getitem__(self, idx):
        return self.materialise()[idx]

    def __len__(self):
        return self.shape[0]

    def __array__(self, *args, **kwargs):         # numpy conversion behaves like the tensor's (a CUDA tensor refuses it)
        return self.materialise().__array__(*args, **kwargs)

    def __repr__(self):
        return f"{type(self).__name__}(shape={tuple(self.shape)}, materialised={self._map is not None})"


class LazyUpsample(_LazyMap):
    """``scale * nn.Upsample(size, mode='bilinear')(src)`` of a detached CUDA heatmap, not yet computed (what the overlay's
    ``nn.Upsample`` route returns for train1.py:410-417).  ``c * lazy`` stays lazy, ``lazy + lazy`` (same output size) becomes
    :class:`FusedHeads` - so the driver's ``target5 = 0.5 * target + target1`` reaches the loss as its two heads."""

    def __init__(self, src, size, scale=1.0):
        self.src = _lib.require_cuda(src.detach(), "LazyUpsample(src)")
        if self.src.ndim != 4:
            raise ValueError("LazyUpsample: src must be [B,K,h,w]")
        self.size = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
        self.scale = float(scale)

    @property
    def shape(self):
        return torch.Size((self.src.shape[0], self.src.shape[1]) + self.size)

    def materialise(self):
        if self._map is None:
            self._map = _fuse(self.src, self.scale, None, 0.0, None, 0.0, self.size)
        return self._map

    def __mul__(self, other):
        if isinstance(other, (int, float)) and not isinstance(other, bool):
            return LazyUpsample(self.src, self.size, self.scale * float(other))
        return self.materialise() * _unlazy(other)

    __rmul__ = __mul__

    def __add__(self, other):
        if (isinstance(other, LazyUpsample) and other.size == self.size and self.size[0] == self.size[1]
                and other.src.shape[:2] == self.src.shape[:2] and other.src.device == self.src.device):
            lo, mid = (self, other) if self.src.shape[2] <= other.src.shape[2] else (other, self)
            return FusedHeads(lo.src, mid.src, lo.scale, mid.scale, self.size[0])
        return self.materialise() + _unlazy(other)

    def __radd__(self, other):
        return self.__add__(other)


class FusedHeads(_LazyMap):
    """``a_lo * up(lo) + a_mid * up(mid)`` at ``size`` x ``size`` - the ``target5`` of train1.py:410-424 - NOT materialised.

    Pass it where ``RegressionDisparityx6.forward`` takes ``y_adv2``
    (``regression_disparity(y_t, y_t_adv, FusedHeads(y_t_adv3, y_t_adv2), weight_t, mode='max')`` instead of building
    ``target5`` first): for the driver's geometry (16x16 and 32x32 heads, 64x64 maps) the loss kernel interpolates the
    fused values from the two heads itself (``hp_regdisp_fwd_heads``; 5 KB read per map instead of 16 KB written and
    16 KB read, no fusion launch), bit-identical to the materialised map.  Every other consumer gets :meth:`materialise`
    (one ``hp_fuse_multiscale`` launch, cached)."""

    def __init__(self, lo, mid, a_lo=0.5, a_mid=1.0, size=64):
        self.lo = _lib.require_cuda(lo.detach(), "FusedHeads(lo)")
        self.mid = _lib.require_cuda(mid.detach(), "FusedHeads(mid)")
        if self.lo.ndim != 4 or self.mid.ndim != 4 or self.mid.shape[:2] != self.lo.shape[:2]:
            raise ValueError("FusedHeads: lo and mid must be [B,K,h,w] with equal batch / joint dims")
        self.a_lo, self.a_mid, self.size = float(a_lo), float(a_mid), int(size)
        self._map = None

    @property
    def shape(self):
        return torch.Size((self.lo.shape[0], self.lo.shape[1], self.size, self.size))

    def in_kernel(self):
        """True when the dense disparity kernel can build the map itself (16x16 / 32x32 -> 64x64)."""
        return (self.size == 64 and tuple(self.lo.shape[2:]) == (16, 16) and tuple(self.mid.shape[2:]) == (32, 32)
                and self.lo.data_ptr() % 16 == 0 and self.mid.data_ptr() % 16 == 0)

    def materialise(self):
        if self._map is None:
            self._map = _fuse(self.lo, self.a_lo, self.mid, self.a_mid, None, 0.0, self.size)
        return self._map
