"""GPU parity of the dense 'max' disparity kernel (csrc/hp_regdisp_dense.cuh: RegressionDisparityx6 / RegressionDisparity4,
regda_7.py:3609-3632, regda_4.py) on the cases its shortcuts have to get right:

  * every other joint's centre INSIDE the own patch (the maximum of the ground-false map is not 1 in closed form: the exact
    per-pixel path), incl. all 21 maps non-positive (all centres (0, 0)) and centres in the map's corners;
  * fused maps that push the ground-false label to all-zero (0 / 0 = NaN in the reference), NaN / +-inf in the fused map
    or the prediction, fused values that make the maximum anything between 0 and 1;
  * few blocks (map ranges over many samples: slot hand-over of the per-sample label) == the full grid, bit for bit;
  * loss ('none' and 'mean'), the saved statistics through the backward kernel (gradient), against the live oracle and
    against the register-slice kernel it replaced (HP_RD_DENSE=0 in a subprocess is not needed: HP_RD_SHAPE=g takes the
    guarded generic kernel).
Tolerances: loss rtol 1e-5; gradient rtol 1e-5 with the absolute slack scaled to the tensor (see test_gpu_parity.py)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import api

pytestmark = pytest.mark.gpu

hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
K = 21


def _peaks(rs, B, centres, amp=1.0, noise=0.02):
    """y[b,k]: a single clear peak at centres[b,k] = (x, y) plus small noise."""
    y = (noise * rs.standard_normal((B, K, 64, 64))).astype(np.float32)
    for b in range(B):
        for k in range(K):
            cx, cy = centres[b, k]
            y[b, k, cy, cx] = amp + rs.uniform(0.1, 0.5)
    return y


def _run(ns, device, y_h, adv_h, f_h, w_h, go_h, mode="max", cls="RegressionDisparityx6"):
    kl = ns.JointsKLLoss(reduction="none", epsilon=1e-7)
    rd = getattr(ns, cls)(ns.PseudoLabelGenerator(K, 64, 64), kl)
    y = torch.from_numpy(y_h).to(device)
    adv = torch.from_numpy(adv_h).to(device).requires_grad_(True)
    w = None if w_h is None else torch.from_numpy(w_h).to(device)
    f = None if f_h is None else torch.from_numpy(f_h).to(device)
    l = rd(y, adv, f, w, mode) if cls == "RegressionDisparityx6" else rd(y, adv, w, mode)
    l.backward(torch.from_numpy(go_h).to(device))
    klm = ns.JointsKLLoss(epsilon=1e-7)
    rdm = getattr(ns, cls)(ns.PseudoLabelGenerator(K, 64, 64), klm)
    with torch.no_grad():
        lm = rdm(y, adv.detach(), f, w, mode) if cls == "RegressionDisparityx6" else rdm(y, adv.detach(), w, mode)
    return l.detach().cpu().numpy(), adv.grad.cpu().numpy(), float(lm)


def _compare(y_h, adv_h, f_h, w_h, seed, cls="RegressionDisparityx6", monkeypatch=None):
    B = y_h.shape[0]
    go_h = np.random.RandomState(seed).uniform(0.5, 1.5, size=(B,)).astype(np.float32)
    ns = hp
    l_gpu, g_gpu, m_gpu = _run(ns, "cuda", y_h, adv_h, f_h, w_h, go_h, cls=cls)
    l_ref, g_ref, m_ref = _run(api.namespace(), "cpu", y_h, adv_h, f_h, w_h, go_h, cls=cls)
    np.testing.assert_allclose(l_gpu, l_ref, rtol=1e-5, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(m_gpu, m_ref, rtol=1e-5, equal_nan=True)
    ok = np.isfinite(g_ref)
    scale = 1e-5 * float(np.abs(g_ref[ok]).max()) if ok.any() else 0.0
    assert np.array_equal(np.isnan(g_gpu), np.isnan(g_ref))
    np.testing.assert_allclose(g_gpu[ok], g_ref[ok], rtol=1e-5, atol=scale)
    if monkeypatch is not None:
        monkeypatch.setenv("HP_RD_GRID", "3")
        l_few, g_few, m_few = _run(ns, "cuda", y_h, adv_h, f_h, w_h, go_h, cls=cls)
        monkeypatch.delenv("HP_RD_GRID")
        assert np.array_equal(l_gpu, l_few, equal_nan=True) and np.array_equal(g_gpu, g_few, equal_nan=True)
        assert m_gpu == m_few or (np.isnan(m_gpu) and np.isnan(m_few))
    return l_gpu


@pytest.mark.parametrize("cls", ["RegressionDisparityx6", "RegressionDisparity4", "RegressionDisparity"])
def test_clustered_centres_take_the_exact_path(cls, monkeypatch):
    """Samples whose 21 decoded centres all lie within each other's 13x13 patches: max(gf) has no closed form."""
    rs = np.random.RandomState(4101)
    B = 9
    centres = np.zeros((B, K, 2), dtype=np.int64)
    centres[0] = rs.randint(30, 34, size=(K, 2))            # tight cluster in the middle
    centres[1] = rs.randint(0, 5, size=(K, 2))              # cluster in the top-left corner (truncated patches)
    centres[2] = rs.randint(59, 64, size=(K, 2))            # bottom-right corner
    centres[3] = np.array([17, 40])                         # all on ONE pixel
    centres[4] = rs.randint(20, 27, size=(K, 2))            # 7x7 cluster: all inside every patch (|d| <= 6)
    centres[5] = rs.randint(0, 64, size=(K, 2))             # ordinary sample (closed-form path)
    centres[6] = rs.randint(20, 27, size=(K, 2))
    centres[6, 20] = (40, 40)                               # one joint far away: maps 0..19 closed form, map 20 exact
    centres[7] = rs.randint(0, 64, size=(K, 2))
    centres[8] = rs.randint(10, 17, size=(K, 2))
    y_h = _peaks(rs, B, centres)
    y_h[7] = -np.abs(y_h[7])                                # every map non-positive: all centres (0, 0)
    adv_h = hp.synth.make_host_batch(4102, B, K, 64, 64)["pred"]
    w_h = (rs.uniform(size=(B, K, 1)) < 0.85).astype(np.float32)
    _compare(y_h, adv_h, None, w_h, 4103, cls=cls, monkeypatch=monkeypatch)


def test_fused_map_edge_values(monkeypatch):
    """x6 'max' with a fused map: maxima anywhere in (0, 1], an all-zero label (0/0 -> NaN like the reference), NaN and
    +-inf in the fused map, a clustered sample, weights None."""
    rs = np.random.RandomState(4201)
    B = 8
    y_h = hp.synth.make_host_batch(4202, B, K, 64, 64)["pred"]
    adv_h = hp.synth.make_host_batch(4203, B, K, 64, 64)["pred"]
    f_h = rs.uniform(-0.3, 0.6, size=(B, K, 64, 64)).astype(np.float32)
    f_h[1] = -3.0                                           # label all zero: NaN losses for the sample
    f_h[2, :, :, :] = rs.uniform(-1.2, -0.6, size=(K, 64, 64)).astype(np.float32)   # max(gf) in (0, 0.4)
    f_h[3, 4, 10, 11] = np.nan
    f_h[3, 5, 0, 0] = np.inf
    f_h[3, 6, 63, 63] = -np.inf
    f_h[4] = 0.0                                            # the fused map adds nothing
    centres = np.tile(rs.randint(28, 34, size=(1, K, 2)), (1, 1, 1))
    y_h[5] = _peaks(rs, 1, centres)[0]                      # clustered centres + fused map
    adv_h[6, 3, 20, 20] = np.nan                            # NaN in the prediction
    l = _compare(y_h, adv_h, f_h, None, 4204, monkeypatch=monkeypatch)
    assert np.isnan(l[1]) and np.isnan(l[3]) and np.isnan(l[6]) and np.isfinite(l[0]) and np.isfinite(l[2])


@pytest.mark.parametrize("fused", [False, True])
def test_large_logits_take_the_exact_softmax(fused, monkeypatch):
    """The passes run the softmax against a SAMPLED reference value; predictions whose maximum lies far above (or whose
    sampled pixels lie far above the rest of) the map overflow / underflow that sum and must come back through the exact
    two-pass path: logits of +-300, a single +200 spike, +inf (NaN like torch), a constant offset of -500 (the sums
    run against the reference, so the offset does not cost precision).  Not covered: -inf logits - torch gives +inf, the
    patch-correction algebra of this kernel (and of the headline pipeline kernel) gives NaN (inf - inf)."""
    rs = np.random.RandomState(4401)
    B = 6
    y_h = hp.synth.make_host_batch(4402, B, K, 64, 64)["pred"]
    adv_h = hp.synth.make_host_batch(4403, B, K, 64, 64)["pred"]
    adv_h[0] *= 300.0
    adv_h[1, :, 40, 17] += 200.0
    adv_h[2, :, :2, :] += 150.0                             # the sampled rows far ABOVE the rest
    adv_h[3] += 500.0
    adv_h[4, 7, 50, 50] = np.inf
    adv_h[5] -= 500.0
    f_h = rs.uniform(-0.3, 0.6, size=(B, K, 64, 64)).astype(np.float32) if fused else None
    l = _compare(y_h, adv_h, f_h, None, 4404, monkeypatch=monkeypatch)
    assert np.isfinite(l[0]) and np.isfinite(l[1]) and np.isfinite(l[2]) and np.isfinite(l[3]) and np.isnan(l[4]) and np.isfinite(l[5])


def test_x5_fused_map_edge_values():
    """x5 'max' with the fused map on the 32x32 head (train1.py:421; one fused decode+loss kernel): maxima anywhere in
    (0, 1], an all-zero label (NaN like the reference), NaN / +-inf in the fused map, NaN in the prediction; loss, 'mean'
    and gradient against the oracle."""
    rs = np.random.RandomState(4501)
    B = 7
    y_h = hp.synth.make_host_batch(4502, B, K, 64, 64)["pred"]
    adv64 = hp.synth.make_host_batch(4503, B, K, 64, 64)["pred"]
    adv_h = hp.synth.make_lowres_heads(4504, adv64, (32,))[0]
    f_h = rs.uniform(-0.9, 0.6, size=(B, K, 32, 32)).astype(np.float32)
    f_h[1] = -3.0
    f_h[2] = rs.uniform(-0.9, -0.5, size=(K, 32, 32)).astype(np.float32)    # max(gf) in (0.1, 0.5)
    f_h[3, 4, 10, 11] = np.nan
    f_h[3, 5, 0, 0] = np.inf
    f_h[3, 6, 31, 31] = -np.inf
    f_h[4] = 0.0
    adv_h[5, 3, 20, 20] = np.nan
    w_h = (rs.uniform(size=(B, K, 1)) < 0.85).astype(np.float32)
    go_h = rs.uniform(0.5, 1.5, size=(B,)).astype(np.float32)

    def run(ns, device):
        rd = ns.RegressionDisparityx5(ns.PseudoLabelGenerator03(K), ns.JointsKLLoss(reduction="none", epsilon=1e-7))
        rdm = ns.RegressionDisparityx5(ns.PseudoLabelGenerator03(K), ns.JointsKLLoss(epsilon=1e-7))
        y, f, w = (torch.from_numpy(a).to(device) for a in (y_h, f_h, w_h))
        adv = torch.from_numpy(adv_h).to(device).requires_grad_(True)
        l = rd(y, adv, f, w, "max")
        l.backward(torch.from_numpy(go_h).to(device))
        with torch.no_grad():
            m = float(rdm(y, adv.detach(), f, w, "max"))
        return l.detach().cpu().numpy(), adv.grad.cpu().numpy(), m

    l_gpu, g_gpu, m_gpu = run(hp, "cuda")
    l_ref, g_ref, m_ref = run(api.namespace(), "cpu")
    np.testing.assert_allclose(l_gpu, l_ref, rtol=1e-5, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(m_gpu, m_ref, rtol=1e-5, equal_nan=True)
    assert np.isnan(l_gpu[1]) and np.isnan(l_gpu[3]) and np.isnan(l_gpu[5]) and np.isfinite(l_gpu[0]) and np.isfinite(l_gpu[2])
    ok = np.isfinite(g_ref)
    assert np.array_equal(np.isnan(g_gpu), np.isnan(g_ref))
    np.testing.assert_allclose(g_gpu[ok], g_ref[ok], rtol=1e-5, atol=1e-5 * float(np.abs(g_ref[ok]).max()))


def test_mean_matches_generic_kernel_at_scale(monkeypatch):
    """150 samples (more than one map range per SM boundary), 'mean': dense kernel == guarded generic kernel to 1e-6."""
    B = 150
    y = torch.from_numpy(hp.synth.make_host_batch(4301, B, K, 64, 64)["pred"]).cuda()
    adv = torch.from_numpy(hp.synth.make_host_batch(4302, B, K, 64, 64)["pred"]).cuda()
    f = torch.from_numpy(np.random.RandomState(4303).uniform(-0.3, 0.6, size=(B, K, 64, 64)).astype(np.float32)).cuda()
    rd = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, 64, 64), hp.JointsKLLoss(epsilon=1e-7))
    rdn = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, 64, 64), hp.JointsKLLoss(reduction="none", epsilon=1e-7))
    with torch.no_grad():
        got = [float(rd(y, adv, ff, None, "max")) for ff in (None, f)]
        got_n = [rdn(y, adv, ff, None, "max").cpu().numpy() for ff in (None, f)]
        monkeypatch.setenv("HP_RD_SHAPE", "g")
        want = [float(rd(y, adv, ff, None, "max")) for ff in (None, f)]
        want_n = [rdn(y, adv, ff, None, "max").cpu().numpy() for ff in (None, f)]
        monkeypatch.delenv("HP_RD_SHAPE")
    np.testing.assert_allclose(got, want, rtol=2e-6)
    for a, b in zip(got_n, want_n):
        np.testing.assert_allclose(a, b, rtol=1e-5)


# ---- x6 'max' with the fused map left UNFUSED (hp.FusedHeads -> hp_regdisp_fwd_heads / _bwd_heads) ---------------------------------

def _heads_case(seed, B):
    rs = np.random.RandomState(seed)
    y_h = hp.synth.make_host_batch(seed + 1, B, K, 64, 64)["pred"]
    adv_h = hp.synth.make_host_batch(seed + 2, B, K, 64, 64)["pred"]
    mid_h, lo_h = hp.synth.make_lowres_heads(seed + 3, adv_h, (32, 16))
    return rs, y_h, adv_h, lo_h, mid_h


def _run_heads(device, y_h, adv_h, lo_h, mid_h, w_h, go_h, how, mode="max"):
    """how: 'heads' (hp.FusedHeads: the map is built inside the loss kernel), 'map' (hp.fuse_multiscale, then the pre-fused
    kernel), 'oracle' (the reference restatement on the CPU: nn.Upsample x 2 + 0.5 a + b, then RegressionDisparityx6)."""
    from oracle import hp_oracle as O
    ns = api.namespace() if how == "oracle" else hp
    t = lambda a: torch.from_numpy(a).to(device)
    if how == "heads":
        f = hp.FusedHeads(t(lo_h), t(mid_h))
    elif how == "map":
        f = hp.fuse_multiscale(t(lo_h), t(mid_h), 64, 32)[0]
    else:
        f = O.fuse_multiscale(t(lo_h), t(mid_h), 64, 32)[0]
    rd = ns.RegressionDisparityx6(ns.PseudoLabelGenerator(K, 64, 64), ns.JointsKLLoss(reduction="none", epsilon=1e-7))
    adv = t(adv_h).requires_grad_(True)
    w = None if w_h is None else t(w_h)
    l = rd(t(y_h), adv, f, w, mode)
    l.backward(t(go_h))
    M = None
    if how != "oracle" and mode == "max":  # the per-map maxima the kernel normalised by (stats[:, 2] of the packed forward output)
        n = y_h.shape[0] * K
        M = rd._lazy[2][n:4 * n].reshape(n, 3)[:, 2].cpu().numpy()
    rdm = ns.RegressionDisparityx6(ns.PseudoLabelGenerator(K, 64, 64), ns.JointsKLLoss(epsilon=1e-7))
    with torch.no_grad():
        m = float(rdm(t(y_h), adv.detach(), f, w, mode))
    return l.detach().cpu().numpy(), adv.grad.cpu().numpy(), m, M


def _close(got, want, rtol):
    l_g, g_g, m_g, _ = got
    l_w, g_w, m_w, _ = want
    np.testing.assert_allclose(l_g, l_w, rtol=rtol, atol=1e-7, equal_nan=True)
    np.testing.assert_allclose(m_g, m_w, rtol=rtol, equal_nan=True)
    ok = np.isfinite(g_w)
    assert np.array_equal(np.isnan(g_g), np.isnan(g_w))
    np.testing.assert_allclose(g_g[ok], g_w[ok], rtol=rtol, atol=rtol * float(np.abs(g_w[ok]).max()) if ok.any() else 0.0)


def test_fused_heads_in_kernel_equals_prefused_map_and_oracle(monkeypatch):
    """train1.py:410-426 with target5 never materialised: the loss kernel interpolates 0.5 up64(y_adv3) + up64(y_adv2) from the
    staged 16x16 / 32x32 heads.  Against (a) the same loss on the materialised map (hp.fuse_multiscale + pre-fused kernel):
    the per-map maxima M are bit-identical (the fused VALUES are the fusion kernel's, bit for bit), losses / gradients agree
    to 2e-6 (only the summation order differs); (b) the oracle, 1e-5; (c) few blocks == the full grid, bit for bit.  Edge
    values: heads that drive the label to all-zero (NaN like the reference), NaN / +-inf inside the heads (incl. the clamped
    first row / column, where a zero-weight tap on inf must still give NaN), clustered centres, NaN in the prediction."""
    rs, y_h, adv_h, lo_h, mid_h = _heads_case(4601, 9)
    B = 9
    mid_h[1] = -3.0                                          # fused map << 0: label all zero -> NaN for the sample
    lo_h[2] = rs.uniform(-1.6, -0.9, size=(K, 16, 16)).astype(np.float32)
    mid_h[2] = rs.uniform(-0.4, 0.1, size=(K, 32, 32)).astype(np.float32)      # max(gf) somewhere in (0, 1)
    lo_h[3, 4, 7, 9] = np.nan
    lo_h[3, 5, 1, 1] = np.inf                                # tap (0, 1) of the clamped first block row / column: 0 * inf
    mid_h[3, 6, 31, 31] = -np.inf
    mid_h[3, 7, 0, 1] = np.inf
    lo_h[4] = 0.0
    mid_h[4] = 0.0                                           # the fused map adds nothing
    centres = rs.randint(28, 34, size=(1, K, 2))
    y_h[5] = _peaks(rs, 1, centres)[0]                       # clustered centres: own patches overlap every other centre
    adv_h[6, 3, 20, 20] = np.nan
    y_h[7] = -np.abs(y_h[7])                                 # all centres (0, 0): patches truncated by the corner
    w_h = (rs.uniform(size=(B, K, 1)) < 0.85).astype(np.float32)
    go_h = rs.uniform(0.5, 1.5, size=(B,)).astype(np.float32)
    heads = _run_heads("cuda", y_h, adv_h, lo_h, mid_h, w_h, go_h, "heads")
    pre = _run_heads("cuda", y_h, adv_h, lo_h, mid_h, w_h, go_h, "map")
    ref = _run_heads("cpu", y_h, adv_h, lo_h, mid_h, w_h, go_h, "oracle")
    assert np.array_equal(heads[3], pre[3], equal_nan=True), "per-map maxima differ: the in-kernel fused values are not the fusion kernel's"
    _close(heads, pre, 2e-6)
    _close(heads, ref, 1e-5)
    l = heads[0]
    assert np.isnan(l[1]) and np.isnan(l[3]) and np.isnan(l[6]) and np.isfinite(l[0]) and np.isfinite(l[2]) and np.isfinite(l[5])
    monkeypatch.setenv("HP_RD_GRID", "3")
    few = _run_heads("cuda", y_h, adv_h, lo_h, mid_h, w_h, go_h, "heads")
    monkeypatch.delenv("HP_RD_GRID")
    assert np.array_equal(heads[0], few[0], equal_nan=True) and np.array_equal(heads[1], few[1], equal_nan=True)
    # weights None, 'mean' through the same kernel
    _close(_run_heads("cuda", y_h[:3], adv_h[:3], lo_h[:3], mid_h[:3], None, go_h[:3], "heads"),
           _run_heads("cpu", y_h[:3], adv_h[:3], lo_h[:3], mid_h[:3], None, go_h[:3], "oracle"), 1e-5)


def test_fused_heads_fall_back_to_the_materialised_map():
    """Where the in-kernel fusion does not apply the unfused argument behaves exactly like the tensor it stands for: mode='min'
    (x6 never reads y_adv2 there), other head sizes (8x8 / 16x16 -> 64x64: hp_fuse_multiscale + the pre-fused kernel), and the
    lazily materialised .ground_false."""
    rs, y_h, adv_h, lo_h, mid_h = _heads_case(4701, 4)
    go_h = rs.uniform(0.5, 1.5, size=(4,)).astype(np.float32)
    a = _run_heads("cuda", y_h, adv_h, lo_h, mid_h, None, go_h, "heads", mode="min")
    b = _run_heads("cuda", y_h, adv_h, lo_h, mid_h, None, go_h, "map", mode="min")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    t = lambda x: torch.from_numpy(x).cuda()
    lo8 = np.ascontiguousarray(lo_h[:, :, ::2, ::2])
    rd = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, 64, 64), hp.JointsKLLoss(reduction="none", epsilon=1e-7))
    fh = hp.FusedHeads(t(lo8), t(lo_h))
    assert not fh.in_kernel()
    l1 = rd(t(y_h), t(adv_h), fh, None, "max")
    gf1 = rd.ground_false.clone()
    l2 = rd(t(y_h), t(adv_h), fh.materialise(), None, "max")
    assert torch.equal(l1, l2) and torch.equal(gf1, rd.ground_false)
    # in-kernel path: .ground_false is built from the materialised map on demand and matches the oracle's
    rd(t(y_h), t(adv_h), hp.FusedHeads(t(lo_h), t(mid_h)), None, "max")
    from oracle import hp_oracle as O
    ns = api.namespace()
    ro = ns.RegressionDisparityx6(ns.PseudoLabelGenerator(K, 64, 64), ns.JointsKLLoss(reduction="none", epsilon=1e-7))
    ro(torch.from_numpy(y_h), torch.from_numpy(adv_h), O.fuse_multiscale(torch.from_numpy(lo_h), torch.from_numpy(mid_h), 64, 32)[0], None, "max")
    np.testing.assert_allclose(rd.ground_false.cpu().numpy(), ro.ground_false.numpy(), rtol=1e-5, atol=1e-6)
