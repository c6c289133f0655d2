"""Host logic of the unbuilt maps the overlay's nn.Upsample route hands out (fusion.LazyUpsample / fusion.FusedHeads;
train1.py:410-428): the driver's own statements - `0.5 * target + target1`, passing `target0` on - must end in the objects
the disparity losses recognise, and ANY other use must see the tensor the eager route would have produced.  CPU only: the
kernel launch behind materialise() is replaced by torch's own bilinear upsample (the CUDA side is tests/test_gpu_dense_disparity.py
and tests/test_train1_overlay.py)."""
import importlib

import pytest
import torch
import torch.nn.functional as F

fusion = importlib.import_module("domain-adaptative-hand-pose-estimation_b200.fusion")


@pytest.fixture()
def cpu_fuse(monkeypatch):
    calls = []

    def fake_fuse(lo, a_lo, mid, a_mid, hi, a_hi, size):
        calls.append((tuple(lo.shape), a_lo, None if mid is None else tuple(mid.shape), a_mid))
        hw = (size, size) if isinstance(size, int) else tuple(size)
        out = a_lo * F.interpolate(lo, size=hw, mode="bilinear", align_corners=False)
        if mid is not None:
            out = out + a_mid * F.interpolate(mid, size=hw, mode="bilinear", align_corners=False)
        return out

    monkeypatch.setattr(fusion, "_fuse", fake_fuse)
    monkeypatch.setattr(fusion._lib, "require_cuda", lambda t, name: t.contiguous())
    return calls


def _heads():
    g = torch.Generator().manual_seed(11)
    return torch.randn(2, 21, 16, 16, generator=g), torch.randn(2, 21, 32, 32, generator=g)


def test_the_drivers_blend_becomes_fused_heads(cpu_fuse):
    a16, a32 = _heads()
    target, target1, target0 = fusion.LazyUpsample(a16, 64), fusion.LazyUpsample(a32, 64), fusion.LazyUpsample(a16, 32)
    target5 = 0.5 * target + target1                        # train1.py:424, verbatim
    assert isinstance(target5, fusion.FusedHeads) and not cpu_fuse, "nothing may be computed yet"
    assert target5.lo is not None and tuple(target5.lo.shape[2:]) == (16, 16) and tuple(target5.mid.shape[2:]) == (32, 32)
    assert (target5.a_lo, target5.a_mid, target5.size) == (0.5, 1.0, 64) and target5.in_kernel()
    assert tuple(target5.shape) == (2, 21, 64, 64) and tuple(target0.shape) == (2, 21, 32, 32)
    swapped = target1 + target * 0.5                        # operand order does not matter: lo is the smaller source
    assert isinstance(swapped, fusion.FusedHeads) and (swapped.a_lo, swapped.a_mid) == (0.5, 1.0)
    want = 0.5 * F.interpolate(a16, size=64, mode="bilinear") + F.interpolate(a32, size=64, mode="bilinear")
    assert torch.allclose(target5.materialise(), want, atol=1e-6) and len(cpu_fuse) == 1
    target5.materialise()
    assert len(cpu_fuse) == 1, "materialise() is cached"
    assert cpu_fuse[0] == ((2, 21, 16, 16), 0.5, (2, 21, 32, 32), 1.0), "one launch for the blend, not three"


def test_every_other_use_sees_the_tensor(cpu_fuse):
    a16, a32 = _heads()
    up = F.interpolate(a16, size=64, mode="bilinear")
    lazy = fusion.LazyUpsample(a16, 64)
    assert torch.allclose(torch.sum(lazy), up.sum(), rtol=1e-5)                     # torch function
    assert torch.allclose(torch.cat([lazy, up], dim=0)[:2], up, atol=1e-6)          # inside a list argument
    assert torch.allclose(up * lazy, up * up, atol=1e-5)                            # Tensor.__mul__(tensor, lazy)
    assert torch.allclose(lazy + up, 2 * up, atol=1e-5) and torch.allclose(up - lazy, torch.zeros_like(up), atol=1e-6)
    assert torch.allclose(lazy.detach().mean(), up.mean(), atol=1e-6)               # tensor attribute / method
    assert lazy.dtype == torch.float32 and lazy.dim() == 4 and len(lazy) == 2
    assert torch.allclose(lazy[0, 3], up[0, 3], atol=1e-6)
    # sums the kernel has no form for are plain tensors: different output sizes, a tensor operand, a non-scalar factor
    other = fusion.LazyUpsample(a32, 128)
    with pytest.raises(RuntimeError):
        lazy + other                                                                # 64x64 + 128x128: torch's own shape error
    assert isinstance(lazy * torch.tensor(2.0), torch.Tensor)
    assert isinstance(fusion.LazyUpsample(a16, 64) * 3, fusion.LazyUpsample)
    fh = 0.5 * fusion.LazyUpsample(a16, 64) + fusion.LazyUpsample(a32, 64)
    assert isinstance(fh + up, torch.Tensor) and isinstance(2.0 * fh, torch.Tensor) and isinstance(-fh, torch.Tensor)
    assert torch.allclose(torch.relu(fh), torch.relu(fh.materialise()))


def test_fused_heads_outside_the_kernels_geometry_materialise(cpu_fuse):
    g = torch.Generator().manual_seed(12)
    a8, a16 = torch.randn(1, 21, 8, 8, generator=g), torch.randn(1, 21, 16, 16, generator=g)
    fh = fusion.LazyUpsample(a8, 64) + fusion.LazyUpsample(a16, 64)
    assert isinstance(fh, fusion.FusedHeads) and not fh.in_kernel()
    want = F.interpolate(a8, size=64, mode="bilinear") + F.interpolate(a16, size=64, mode="bilinear")
    assert torch.allclose(fh.materialise(), want, atol=1e-6)


def test_fuse_multiscale_lazy_form(cpu_fuse):
    a16, a32 = _heads()
    t5, t0 = fusion.fuse_multiscale(a16, a32, 64, 32, lazy=True)
    assert isinstance(t5, fusion.FusedHeads) and isinstance(t0, fusion.LazyUpsample) and not cpu_fuse
    assert t5.in_kernel() and tuple(t0.shape) == (2, 21, 32, 32)
    assert torch.allclose(t0.materialise(), F.interpolate(a16, size=32, mode="bilinear"), atol=1e-6)
