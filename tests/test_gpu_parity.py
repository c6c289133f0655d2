"""GPU parity: the CUDA path (through the C ABI) against (a) the golden vectors frozen from the real
reference and (b) the live CPU oracle on the same seeded inputs.  Bars are in the key prefixes of
oracle/cases.py: ``x:`` bit-exact (coordinates, maxvals, PCK counts/acc, weights), ``c:`` rtol 1e-5
(+ atol 1e-6 for heatmaps, scaled for gradients)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import api, cases
from oracle import hp_oracle as O
from oracle.gen_golden import digest

pytestmark = pytest.mark.gpu

hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")


def _ns():
    import types
    return types.SimpleNamespace(
        get_max_preds=hp.get_max_preds, accuracy=hp.accuracy, generate_target=hp.generate_target,
        find_keypoints_max=hp.find_keypoints_max, compute_uv_from_heatmaps=hp.compute_uv_from_heatmaps,
        compute_uv_from_heatmaps2=hp.compute_uv_from_heatmaps2, compute_uv_from_heatmaps3=hp.compute_uv_from_heatmaps3,
        JointsMSELoss=hp.JointsMSELoss, JointsKLLoss=hp.JointsKLLoss,
        PseudoLabelGenerator=hp.PseudoLabelGenerator, PseudoLabelGenerator01=hp.PseudoLabelGenerator01,
        PseudoLabelGenerator02=hp.PseudoLabelGenerator02, PseudoLabelGenerator03=hp.PseudoLabelGenerator03,
        RegressionDisparity=hp.RegressionDisparity, RegressionDisparityx1=hp.RegressionDisparityx1,
        RegressionDisparityx5=hp.RegressionDisparityx5, RegressionDisparityx6=hp.RegressionDisparityx6,
        fuse_multiscale=hp.fuse_multiscale, FusedHeads=hp.FusedHeads, upsample_bilinear=hp.upsample_bilinear,
        **{n: getattr(hp, n) for n in ("RegressionDisparity2", "RegressionDisparity3", "RegressionDisparity4",
                                       "RegressionDisparity5", "RegressionDisparity6", "RegressionDisparity7",
                                       "RegressionDisparity8", "RegressionDisparityx2", "RegressionDisparityx3",
                                       "RegressionDisparityx4", "JointsMSELoss0", "JointsKLLoss5")})


@pytest.fixture(scope="module")
def cuda_outputs():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = cases.CASES[name](_ns(), "cuda")
        return cache[name]
    return get


@pytest.mark.parametrize("name", list(cases.CASES))
def test_cuda_matches_reference_goldens(name, golden, cuda_outputs):
    cases.compare(digest(cuda_outputs(name)), golden(name))


@pytest.mark.parametrize("name", list(cases.CASES))
def test_cuda_matches_live_oracle(name, cuda_outputs):
    want = cases.CASES[name](api.namespace(), "cpu")
    cases.compare(cuda_outputs(name), want)


def test_generated_targets_are_bit_equal_to_live_oracle():
    """Same host, same numpy: the table trick makes generated heatmaps bit-identical, not just close."""
    joints, vis = cases.target_inputs()
    for s in (64, 32, 128):
        t, w = hp.generate_target_batch(joints, vis, (s, s), 2, (256, 256))
        ot, ow = O.generate_target_batch(joints, vis, (s, s), 2, (256, 256))
        assert np.array_equal(t.cpu().numpy(), ot)
        assert np.array_equal(w.cpu().numpy(), ow)


# ------------------------------------------------------------------ fused pipeline

def _run_pipeline(d, eps, **kw):
    pipe = hp.HeatmapPipeline(num_keypoints=d["pred"].shape[1], heatmap_size=(d["pred"].shape[3], d["pred"].shape[2]),
                              kl_epsilon=eps, **kw)
    out = pipe(torch.from_numpy(d["pred"]).cuda(), torch.from_numpy(d["joints"]).cuda(),
               torch.from_numpy(d["vis"]).cuda())
    return pipe, out


@pytest.mark.parametrize("size,B,seed", [(64, 8, 801), (32, 6, 802), (16, 6, 803), (128, 2, 804)])
def test_pipeline_matches_oracle(size, B, seed):
    d = hp.synth.make_host_batch(seed, B, 21, size, size)
    # edge rows: out-of-bounds joint, invisible joint, joint whose centre has x<=1 (PCK-invalid)
    d["joints"][0, 0] = (-50.0, 10.0); d["joints"][0, 1] = (3.0, 100.0); d["vis"][0, 2] = 0.0
    want = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
    _, out = _run_pipeline(d, 1e-7)
    got = out.host()
    part = out.partial.cpu().numpy()
    assert np.array_equal(out.pred_xy.cpu().numpy(), want["pred_xy"])
    assert np.array_equal(out.weight.cpu().numpy(), want["weight"])
    assert np.array_equal(part[4:25].astype(np.int64), want["hits"])
    assert np.array_equal(part[25:46].astype(np.int64), want["valid"])
    assert np.array_equal(got["acc"], want["acc"]) and got["avg_acc"] == want["avg_acc"] and got["cnt"] == want["cnt"]
    np.testing.assert_allclose(got["mse"], want["mse"], rtol=1e-5)
    np.testing.assert_allclose(got["kl"], want["kl"], rtol=1e-5)
    assert part.dtype == np.int64 and part[2] == B * 21 and part[3] == B * 21 * size * size


def test_pipeline_kl_eps0_is_nan_like_reference():
    """epsilon=0 + an invisible joint (all-zero target) -> NaN even with weight 0."""
    d = hp.synth.make_host_batch(811, 4)
    d["vis"][1, 3] = 0.0
    want = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=0.0)
    _, out = _run_pipeline(d, 0.0)
    got = out.host()
    assert np.isnan(want["kl"]) and np.isnan(got["kl"])
    np.testing.assert_allclose(got["mse"], want["mse"], rtol=1e-5)


def test_pipeline_odd_sizes_scalar_walk():
    d = hp.synth.make_host_batch(821, 3, 7, 30, 42)          # H=30, W=42: W % 4 != 0
    want = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
    _, out = _run_pipeline(d, 1e-7)
    got = out.host()
    assert np.array_equal(out.pred_xy.cpu().numpy(), want["pred_xy"])
    assert np.array_equal(got["acc"], want["acc"]) and got["cnt"] == want["cnt"]
    np.testing.assert_allclose(got["mse"], want["mse"], rtol=1e-5)
    np.testing.assert_allclose(got["kl"], want["kl"], rtol=1e-5)


def test_pipeline_host_path_equals_device_path():
    d = hp.synth.make_host_batch(831, 37)
    pipe, out = _run_pipeline(d, 1e-7)
    dev = out.host()
    for slab in (8, 16, 64):
        h = pipe.run_host(d["pred"], d["joints"], d["vis"], slab=slab)
        assert np.array_equal(h["pred_xy"], out.pred_xy.cpu().numpy())
        assert np.array_equal(h["acc"], dev["acc"]) and h["cnt"] == dev["cnt"] and h["avg_acc"] == dev["avg_acc"]
        np.testing.assert_allclose(h["mse"], dev["mse"], rtol=1e-12)
        np.testing.assert_allclose(h["kl"], dev["kl"], rtol=1e-12)


def test_pipeline_full_size_properties():
    """BASELINE configs[1] size (256x21x64x64): size-independent properties instead of the slow oracle -
    fused == composition of the unfused CUDA ops; PCK of targets against themselves; determinism."""
    B = 256
    d = hp.synth.make_device_batch(841, B)
    pipe = hp.HeatmapPipeline(kl_epsilon=1e-7)
    out = pipe(d["pred"], d["joints"], d["vis"])
    r1 = out.result.clone(); p1 = out.partial.clone(); xy1 = out.pred_xy.clone()
    target, weight = hp.generate_target_batch(d["joints"], d["vis"], (64, 64), 2, (256, 256))
    mse = hp.JointsMSELoss()(d["pred"], target, weight)
    kl = hp.JointsKLLoss(epsilon=1e-7)(d["pred"], target, weight)
    acc, avg, cnt, pred = hp.accuracy(d["pred"], target)
    got = out.host()
    assert torch.equal(pred, out.pred_xy) and torch.equal(weight, out.weight)
    assert np.array_equal(acc, got["acc"]) and avg == got["avg_acc"] and cnt == got["cnt"]
    np.testing.assert_allclose(got["mse"], mse.item(), rtol=1e-5)
    np.testing.assert_allclose(got["kl"], kl.item(), rtol=1e-5)
    # decode(generate_target(j)) is the centre: every valid joint hits when scored against itself
    acc_t, avg_t, cnt_t, _ = hp.accuracy(target, target)
    assert all(a in (1.0, -1.0) for a in acc_t)
    # bitwise determinism over repeated launches (fixed-order reductions, integer counters)
    for _ in range(3):
        out2 = pipe(d["pred"], d["joints"], d["vis"])
        assert torch.equal(out2.result, r1) and torch.equal(out2.partial, p1) and torch.equal(out2.pred_xy, xy1)


def _assert_pipeline_equals_oracle(d, out, want, K):
    got = out.host()
    part = out.partial.cpu().numpy()
    assert np.array_equal(out.pred_xy.cpu().numpy(), want["pred_xy"])
    _, want_max = O.get_max_preds(d["pred"])
    assert np.array_equal(out.maxvals.cpu().numpy().reshape(-1), want_max.reshape(-1), equal_nan=True)
    assert np.array_equal(out.weight.cpu().numpy(), want["weight"])
    assert np.array_equal(part[4:4 + K].astype(np.int64), want["hits"])
    assert np.array_equal(part[4 + K:4 + 2 * K].astype(np.int64), want["valid"])
    assert np.array_equal(got["acc"], want["acc"]) and got["avg_acc"] == want["avg_acc"] and got["cnt"] == want["cnt"]
    return got


@pytest.mark.parametrize("size", [64, 32, 16, 128])
def test_pipeline_argmax_edge_maps(size):
    """Ties in every (iteration, lane, component) position, all<=0, -inf, +inf and NaN maps through the
    TMA-staged kernel: coordinates / maxvals / PCK bit-exact against the oracle (numpy: first index wins,
    NaN beats everything and is masked to (0,0))."""
    B, K = 6, 21
    d = hp.synth.make_host_batch(861 + size, B, K, size, size)
    rs = np.random.RandomState(5)
    flat = d["pred"].reshape(B * K, -1)
    HW = size * size
    for m in range(0, 60):                       # duplicated maxima at random positions
        pos = np.sort(rs.choice(HW, size=3, replace=False))
        flat[m, pos] = flat[m].max() + 1.0
    flat[60, :] = -1.0                           # all <= 0 -> (0,0)
    flat[62, :] = 0.25; flat[62, HW - 1] = 0.25  # constant map: index 0
    flat[66, HW - 1] = flat[66].max() + 2.0      # maximum in the very last element
    flat[67, 0] = flat[67].max() + 2.0           # and in the very first
    want = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
    _, out = _run_pipeline(d, 1e-7)
    got = _assert_pipeline_equals_oracle(d, out, want, K)
    np.testing.assert_allclose(got["mse"], want["mse"], rtol=1e-5)
    np.testing.assert_allclose(got["kl"], want["kl"], rtol=1e-5)
    # non-finite maps: the losses become NaN / inf exactly like the reference's
    flat[61, :] = -np.inf                        # all -inf
    flat[63, rs.randint(HW)] = np.inf            # +inf
    flat[64, HW // 2 + 3] = np.nan               # one NaN
    flat[65, 5] = np.nan; flat[65, 4] = np.inf   # NaN after a larger value: NaN still wins
    want = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
    _, out = _run_pipeline(d, 1e-7)
    got = _assert_pipeline_equals_oracle(d, out, want, K)
    for name in ("mse", "kl"):
        assert not np.isfinite(want[name])
        assert np.isnan(got[name]) == np.isnan(want[name]) and (np.isnan(want[name]) or got[name] == want[name])


def test_pipeline_many_maps_per_warp():
    """More than 32 maps per warp (second batch of lane-distributed keypoints) and a wrapped stage ring:
    2800 x 21 maps of 16x16 and 40 x 21 maps of 128x128 against the oracle."""
    for size, B, seed in ((16, 2800, 871), (128, 40, 872)):
        d = hp.synth.make_host_batch(seed, B, 21, size, size)
        want = O.pipeline(d["pred"], d["joints"], d["vis"], kl_epsilon=1e-7)
        _, out = _run_pipeline(d, 1e-7)
        got = _assert_pipeline_equals_oracle(d, out, want, 21)
        np.testing.assert_allclose(got["mse"], want["mse"], rtol=1e-5)
        np.testing.assert_allclose(got["kl"], want["kl"], rtol=1e-5)


def test_pipeline_overlapped_launches_are_bit_identical():
    """HP_PIPE_OVERLAP_PREV (programmatic dependent launch): a train of back-to-back launches over resident
    batches with separate outputs gives exactly the results of fully serialised launches."""
    B, n = 256, 6
    sets = [hp.synth.make_device_batch(881 + i, B) for i in range(n)]
    pipe = hp.HeatmapPipeline(kl_epsilon=1e-7)
    ref = []
    for s in sets:
        o = pipe(s["pred"], s["joints"], s["vis"])
        ref.append((o.result.clone(), o.partial.clone(), o.pred_xy.clone(), o.maxvals.clone(), o.weight.clone()))
    for depth in (True, 1, 2, 3, 8):
        outs = [pipe.alloc_outputs(B) for _ in range(n)]
        for rep in range(5):
            for s, o in zip(sets, outs):
                pipe(s["pred"], s["joints"], s["vis"], out=o, overlap=depth)
        torch.cuda.synchronize()
        for o, r in zip(outs, ref):
            assert torch.equal(o.result, r[0]) and torch.equal(o.partial, r[1]) and torch.equal(o.pred_xy, r[2])
            assert torch.equal(o.maxvals, r[3]) and torch.equal(o.weight, r[4])
    # a small batch (fewer maps than block slots) and a big one (outputs beyond the shared-memory buffer)
    for Bx, size in ((3, 64), (1200, 16)):
        d = hp.synth.make_device_batch(891 + Bx, Bx, 21, size, size)
        p2 = hp.HeatmapPipeline(heatmap_size=(size, size), kl_epsilon=1e-7)
        o = p2(d["pred"], d["joints"], d["vis"])
        want = (o.result.clone(), o.partial.clone(), o.pred_xy.clone(), o.maxvals.clone(), o.weight.clone())
        outs = [p2.alloc_outputs(Bx) for _ in range(4)]
        for rep in range(3):
            for o in outs:
                p2(d["pred"], d["joints"], d["vis"], out=o, overlap=True)
        torch.cuda.synchronize()
        for o in outs:
            assert torch.equal(o.result, want[0]) and torch.equal(o.partial, want[1]) and torch.equal(o.pred_xy, want[2])
            assert torch.equal(o.maxvals, want[3]) and torch.equal(o.weight, want[4])


# ------------------------------------------------------------------ generic shapes / edge cases

@pytest.mark.parametrize("H,W,K", [(30, 42, 5), (20, 20, 21), (100, 100, 3), (7, 5, 2), (96, 72, 4), (256, 256, 2)])
def test_decode_and_accuracy_generic_shapes(H, W, K):
    d = hp.synth.make_host_batch(900 + H, 3, K, H, W)
    p, mv = hp.get_max_preds(d["pred"])
    op, omv = O.get_max_preds(d["pred"])
    assert np.array_equal(p, op) and np.array_equal(mv, omv)
    tgt, _ = O.generate_target_batch(d["joints"], d["vis"], (W, H), 2, (256, 256))
    acc, avg, cnt, pred = hp.accuracy(d["pred"], tgt)
    oacc, oavg, ocnt, opred = O.accuracy(d["pred"], tgt)
    assert np.array_equal(acc, oacc) and avg == oavg and cnt == ocnt and np.array_equal(pred, opred)


def test_pck_threshold_boundaries_non_power_of_two():
    """dx,dy sweeps at sizes where 0.5*norm lands on representable distances (20, 40, 100, 50)."""
    for S in (20, 40, 50, 100):
        K = 21
        o = np.zeros((4, K, S, S), np.float32); t = np.zeros((4, K, S, S), np.float32)
        rs = np.random.RandomState(S)
        for b in range(4):
            for k in range(K):
                cx, cy = rs.randint(8, S - 8, size=2)
                dx, dy = rs.randint(-6, 7, size=2)
                t[b, k, cy, cx] = 1.0
                o[b, k, cy + dy, cx + dx] = 1.0
        acc, avg, cnt, _ = hp.accuracy(o, t)
        oacc, oavg, ocnt, _ = O.accuracy(o, t)
        assert np.array_equal(acc, oacc) and avg == oavg and cnt == ocnt, S


@pytest.mark.parametrize("H,W", [(30, 42), (48, 48), (24, 36)])
def test_losses_generic_shapes(H, W):
    d = hp.synth.make_host_batch(950 + H, 3, 6, H, W)
    tgt, w = O.generate_target_batch(d["joints"], d["vis"], (W, H), 2, (256, 256))
    for make_g, make_o in ((lambda: hp.JointsMSELoss(), lambda: api.JointsMSELoss()),
                           (lambda: hp.JointsKLLoss(epsilon=1e-7), lambda: api.JointsKLLoss(epsilon=1e-7)),
                           (lambda: hp.JointsKLLoss("none", 1e-7), lambda: api.JointsKLLoss("none", 1e-7)),
                           (lambda: hp.JointsMSELoss("none"), lambda: api.JointsMSELoss("none"))):
        pg = torch.from_numpy(d["pred"]).cuda().requires_grad_(True)
        pc = torch.from_numpy(d["pred"]).requires_grad_(True)
        lg = make_g()(pg, torch.from_numpy(tgt).cuda(), torch.from_numpy(w).cuda())
        lc = make_o()(pc, torch.from_numpy(tgt), torch.from_numpy(w))
        np.testing.assert_allclose(lg.detach().cpu().numpy(), lc.detach().numpy(), rtol=1e-5, atol=1e-7)
        lg.sum().backward(); lc.sum().backward()
        g = pc.grad.numpy()
        np.testing.assert_allclose(pg.grad.cpu().numpy(), g, rtol=1e-5, atol=1e-6 * np.abs(g).max())


def test_three_scale_fuse_decode_pck_matches_unfused_and_oracle():
    rs = np.random.RandomState(77)
    B, K = 4, 21
    d = hp.synth.make_host_batch(78, B, K, 128, 128)
    mid, lo = hp.synth.make_lowres_heads(79, d["pred"], (64, 32))
    hi = (0.3 * d["pred"]).astype(np.float32)
    tgt, _ = O.generate_target_batch(d["joints"], d["vis"], (128, 128), 2, (256, 256))
    txy, _ = O.get_max_preds(tgt)
    ev = hp.MultiscaleEval(K)
    acc, pred_xy, counts = ev(torch.from_numpy(lo).cuda(), torch.from_numpy(mid).cuda(), torch.from_numpy(hi).cuda(),
                              torch.from_numpy(txy).cuda())
    fused_gpu = hp.fuse_three_scales(torch.from_numpy(lo).cuda(), torch.from_numpy(mid).cuda(), torch.from_numpy(hi).cuda())
    fused_cpu = O.fuse_three_scales(torch.from_numpy(lo), torch.from_numpy(mid), torch.from_numpy(hi))
    np.testing.assert_allclose(fused_gpu.cpu().numpy(), fused_cpu.numpy(), rtol=1e-5, atol=1e-6)
    # in-register fusion decodes exactly like decoding the materialised CUDA fusion
    p2, _ = hp.get_max_preds(fused_gpu)
    assert torch.equal(pred_xy, p2)
    hits, valid = O.pck_counts(p2.cpu().numpy(), txy, 128, 128)
    c = counts.cpu().numpy()
    assert np.array_equal(c[:K], hits) and np.array_equal(c[K:], valid)


@pytest.mark.parametrize("variant,mode,fused", [
    ("base", "min", False), ("base", "max", False), ("x1", "min", False), ("x1", "max", False),
    ("x5", "min", False), ("x5", "max", False), ("x5", "max", True),
    ("x6", "min", False), ("x6", "max", False), ("x6", "max", True)])
def test_disparity_long_map_ranges(variant, mode, fused, monkeypatch):
    """The staged disparity kernels give each block a contiguous range of maps.  B=2 cases keep every range to
    one map; here 17 samples run on the full grid AND on 5 blocks (ranges of ~71 maps across 4-5 samples:
    sample hand-over, ring wrap-around) and must agree bit-for-bit with each other, to 1e-5 with the guarded
    generic kernel (HP_RD_SHAPE=g) and with the live oracle, for the loss ('none' reduction) and the gradient."""
    B = 17
    d = hp.synth.make_host_batch(911, B, 21, 64, 64)
    adv64 = hp.synth.make_host_batch(912, B, 21, 64, 64)["pred"]
    adv32, adv16 = hp.synth.make_lowres_heads(913, adv64, (32, 16))
    rs = np.random.RandomState(914)
    w_h = (rs.uniform(size=(B, 21, 1)) < 0.85).astype(np.float32)
    side = {"base": 64, "x1": 16, "x5": 32, "x6": 64}[variant]
    adv_h = {64: adv64, 32: adv32, 16: adv16}[side]
    f_h = np.clip(hp.synth.make_host_batch(915, B, 21, side, side)["pred"], -0.2, 1.2).astype(np.float32) if fused else None
    go_h = rs.uniform(0.5, 1.5, size=(B,)).astype(np.float32)

    def run(ns, device):
        kl = ns.JointsKLLoss(reduction="none", epsilon=1e-7)
        rd = {"base": lambda: ns.RegressionDisparity(ns.PseudoLabelGenerator(21, 64, 64), kl),
              "x1": lambda: ns.RegressionDisparityx1(ns.PseudoLabelGenerator01(21), kl),
              "x5": lambda: ns.RegressionDisparityx5(ns.PseudoLabelGenerator03(21), kl),
              "x6": lambda: ns.RegressionDisparityx6(ns.PseudoLabelGenerator(21, 64, 64), kl)}[variant]()
        y = torch.from_numpy(d["pred"]).to(device)
        adv = torch.from_numpy(adv_h).to(device).requires_grad_(True)
        w = torch.from_numpy(w_h).to(device)
        f = None if f_h is None else torch.from_numpy(f_h).to(device)
        l = rd(y, adv, w, mode) if variant in ("base", "x1") else rd(y, adv, f, w, mode)
        l.backward(torch.from_numpy(go_h).to(device))
        return l.detach().cpu().numpy(), adv.grad.cpu().numpy()

    ns = _ns()
    l_full, g_full = run(ns, "cuda")
    monkeypatch.setenv("HP_RD_GRID", "5")
    l_few, g_few = run(ns, "cuda")
    monkeypatch.delenv("HP_RD_GRID")
    monkeypatch.setenv("HP_RD_SHAPE", "g")
    l_gen, g_gen = run(ns, "cuda")
    monkeypatch.delenv("HP_RD_SHAPE")
    l_ref, g_ref = run(api.namespace(), "cpu")
    assert np.array_equal(l_full, l_few) and np.array_equal(g_full, g_few), "results depend on the grid"
    scale = 1e-5 * float(np.abs(g_ref).max())
    for l_o, g_o, what in ((l_gen, g_gen, "generic kernel"), (l_ref, g_ref, "oracle")):
        np.testing.assert_allclose(l_full, l_o, rtol=1e-5, err_msg=what)
        np.testing.assert_allclose(g_full, g_o, rtol=1e-5, atol=scale, err_msg=what)


@pytest.mark.parametrize("sizes", [(16, 32, 64), (32, 64, 128), (8, 16, 32)])
def test_block_fusion_kernel_equals_row_walking_kernels_bit_for_bit(sizes, monkeypatch):
    """The static-pattern kernel for exact x2 / x4 scales (hp_fusion_block.cuh) against the row-walking kernels
    (HP_FUSE_SHAPE=r) and the oracle: materialised maps bit-identical (same arithmetic, incl. the clamped first / last
    taps where 0 * inf must still be NaN), decoded coordinates and PCK counts identical, on inputs that hold NaN, +inf and
    -inf, an all -inf map and an exact tie."""
    lo_s, mid_s, hi_s = sizes
    B, K = 3, 7
    rs = np.random.RandomState(4100 + hi_s)
    hi = rs.standard_normal((B, K, hi_s, hi_s)).astype(np.float32)
    mid = rs.standard_normal((B, K, mid_s, mid_s)).astype(np.float32)
    lo = rs.standard_normal((B, K, lo_s, lo_s)).astype(np.float32)
    lo[0, 1, 1, :] = np.inf                 # 0 * inf at the clamped first row / column
    mid[0, 2, -1, -1] = -np.inf
    mid[0, 3, 0, 1] = np.nan
    hi[1, 0] = -np.inf                      # all -inf map -> element 0
    lo[1, 0] = 0.0
    mid[1, 0] = 0.0
    hi[1, 1], lo[1, 1], mid[1, 1] = 0.0, 0.0, 0.0
    hi[1, 1, 5, 9] = hi[1, 1, 5, 3] = 2.0   # exact tie -> lower index
    tgt = rs.randint(0, hi_s, size=(B, K, 2)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).cuda()

    def run():
        ev = hp.MultiscaleEval(K)
        acc, pred_xy, counts = ev(t(lo), t(mid), t(hi), t(tgt))
        return dict(three=hp.fuse_three_scales(t(lo), t(mid), t(hi)).cpu().numpy(),
                    two=hp.fuse_multiscale(t(lo), t(mid), hi_s, mid_s)[0].cpu().numpy(),
                    up2=hp.upsample_bilinear(t(lo), mid_s).cpu().numpy(),
                    up4=hp.upsample_bilinear(t(lo), hi_s).cpu().numpy(),
                    pred=pred_xy.cpu().numpy(), counts=counts.cpu().numpy(), acc=acc.cpu().numpy())

    got = run()
    monkeypatch.setenv("HP_FUSE_SHAPE", "r")
    ref = run()
    monkeypatch.delenv("HP_FUSE_SHAPE")
    for k in got:
        assert np.array_equal(got[k], ref[k], equal_nan=True), f"{k}: block kernel differs from the row-walking kernel"
    want = O.fuse_three_scales(torch.from_numpy(lo), torch.from_numpy(mid), torch.from_numpy(hi)).numpy()
    finite = np.isfinite(want) & np.isfinite(got["three"])
    assert np.array_equal(np.isnan(want), np.isnan(got["three"]))
    np.testing.assert_allclose(got["three"][finite], want[finite], rtol=1e-5, atol=1e-6)
    p_o, _ = O.get_max_preds(got["three"])
    assert np.array_equal(got["pred"], p_o)


@pytest.mark.parametrize("B", [3, 256])
def test_staged_loss_and_accuracy_kernels_equal_block_kernels(B, monkeypatch):
    """64x64 maps take the warp-private copy-engine kernels (hp_loss_staged.cuh, hp_decode_staged.cuh); the block-per-map
    kernels (HP_LOSS_SHAPE=b / HP_ACC_SHAPE=b) and the oracle are the witnesses for both staged shapes (one warp or a
    warp pair per stage): decoded coordinates, PCK counts and
    accuracies bit-identical, losses / gradients 1e-5 - at the BASELINE.json batch (warps own 6-7 maps each, the
    stage phases alternate) and on a tiny batch, on inputs holding NaN, +-inf, an all -inf map, exact ties, an all-zero
    target and zero weights."""
    K, S = 21, 64
    d = hp.synth.make_host_batch(7300 + B, B, K, S, S, image_size=4 * S)
    pred = d["pred"].copy()
    tg = hp.generate_target_batch(torch.from_numpy(d["joints"]).cuda(), torch.from_numpy(d["vis"]).cuda(), (S, S), 2,
                                  (4 * S, 4 * S))
    tgt = np.ascontiguousarray(tg[0].cpu().numpy(), dtype=np.float32).reshape(B, K, S, S)
    w = np.ascontiguousarray(tg[1].cpu().numpy(), dtype=np.float32).reshape(B, K, 1)
    pred_edge = pred.copy()
    pred_edge[0, 0, 3, 5] = np.nan
    pred_edge[0, 1, 7, 7] = np.inf
    pred_edge[0, 2, 1, 1], pred_edge[0, 2, 9, 9] = np.inf, -np.inf
    pred_edge[0, 3] = -np.inf
    pred_edge[0, 4] = 0.0
    pred_edge[0, 4, 5, 9] = pred_edge[0, 4, 5, 3] = 2.0
    pred_edge[0, 5] = -0.0
    pred_edge[0, 5, 2, 2] = 0.0
    t = lambda a: torch.from_numpy(a).cuda()

    def run(p_np):
        out = {}
        acc, avg, cnt, xy = hp.accuracy(t(p_np), t(tgt))
        out["acc"], out["avg"], out["cnt"], out["xy"] = acc, np.float64(avg), np.int64(cnt), xy.cpu().numpy()
        for name, crit in (("mse", hp.JointsMSELoss()), ("mse_none", hp.JointsMSELoss("none")),
                           ("kl", hp.JointsKLLoss(epsilon=1e-7)), ("kl_none", hp.JointsKLLoss("none", 1e-7)),
                           ("kl_eps0", hp.JointsKLLoss(epsilon=0.0))):
            x = t(p_np).requires_grad_(True)
            l = crit(x, t(tgt), t(w))
            out[name] = l.detach().cpu().numpy()
            if p_np is pred:
                l.sum().backward()
                out[name + "_grad"] = x.grad.cpu().numpy()
        out["kl_now"] = hp.JointsKLLoss(epsilon=1e-7)(t(p_np), t(tgt)).cpu().numpy()
        return out

    for p_np, what0 in ((pred, "synthetic"), (pred_edge, "edge maps")):
        monkeypatch.setenv("HP_LOSS_SHAPE", "b")
        monkeypatch.setenv("HP_ACC_SHAPE", "b")
        ref = run(p_np)
        runs = {}
        for shape in ("1", "2"):        # one warp / a warp pair per stage, for every operator
            monkeypatch.setenv("HP_LOSS_SHAPE", shape)
            monkeypatch.setenv("HP_ACC_SHAPE", shape)
            runs[f"{shape} warp(s) per stage"] = run(p_np)
        monkeypatch.delenv("HP_LOSS_SHAPE")
        monkeypatch.delenv("HP_ACC_SHAPE")
        runs["production shapes"] = got = run(p_np)
        for shape, res in runs.items():
            what = f"{what0}, {shape}"
            for k in ("acc", "avg", "cnt", "xy"):
                assert np.array_equal(res[k], ref[k], equal_nan=True), f"{what}: {k} differs between the staged and block kernels"
            for k in res:
                if k in ("acc", "avg", "cnt", "xy"):
                    continue
                assert np.array_equal(np.isnan(res[k]), np.isnan(ref[k])), f"{what}: {k} NaN pattern"
                scale = 1e-9 + (1e-6 * np.nanmax(np.abs(ref[k])) if k.endswith("_grad") else 0.0)
                np.testing.assert_allclose(res[k], ref[k], rtol=1e-5, atol=scale, equal_nan=True, err_msg=f"{what}: {k}")
        what = what0
        # the oracle on the same inputs
        a_o, avg_o, cnt_o, xy_o = O.accuracy(p_np, tgt)
        assert np.array_equal(got["xy"], xy_o) and np.array_equal(got["acc"], a_o) and got["cnt"] == cnt_o, what
        kl_o = O.joints_kl_loss(torch.from_numpy(p_np), torch.from_numpy(tgt), torch.from_numpy(w), "mean", 1e-7).numpy()
        mse_o = O.joints_mse_loss(torch.from_numpy(p_np), torch.from_numpy(tgt), torch.from_numpy(w), "mean").numpy()
        np.testing.assert_allclose(got["kl"], kl_o, rtol=1e-5, equal_nan=True, err_msg=what)
        np.testing.assert_allclose(got["mse"], mse_o, rtol=1e-5, equal_nan=True, err_msg=what)


def test_foreign_criterion_gets_materialised_maps():
    I = cases.disparity_inputs()
    y, adv, w = (torch.from_numpy(I[k]).cuda() for k in ("y", "adv64", "w"))
    rd_g = hp.RegressionDisparity(hp.PseudoLabelGenerator(21, 64, 64), hp.JointsMSELoss())
    rd_o = api.RegressionDisparity(api.PseudoLabelGenerator(21, 64, 64), api.JointsMSELoss())
    for mode in ("min", "max"):
        lg = rd_g(y, adv, w, mode)
        lo = rd_o(torch.from_numpy(I["y"]), torch.from_numpy(I["adv64"]), torch.from_numpy(I["w"]), mode)
        np.testing.assert_allclose(lg.item(), lo.item(), rtol=1e-5)


def test_no_cpu_fallback_and_argument_errors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hp.JointsKLLoss()(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(AssertionError):
        hp.get_max_preds(np.zeros((2, 4, 4), np.float32))
    with pytest.raises(AssertionError):
        hp.get_max_preds([[1.0]])
    with pytest.raises(ValueError):
        hp.JointsMSELoss()(torch.zeros(1, 2, 4, 4).cuda(), torch.zeros(1, 2, 4, 5).cuda())
    with pytest.raises(IndexError):
        hp.PseudoLabelGenerator01(21)(torch.zeros(1, 21, 128, 128).cuda())
    with pytest.raises(AssertionError):
        hp.RegressionDisparity(hp.PseudoLabelGenerator(21), hp.JointsKLLoss())(
            torch.zeros(1, 21, 64, 64).cuda(), torch.zeros(1, 21, 64, 64).cuda(), None, "sideways")


def test_sharded_pipeline_two_gpus():
    """N>1 on real GPUs (skipped on a single-GPU box): torchrun 2 ranks, NCCL all-reduce of the partials."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29633",
                        os.path.join(root, "tests", "dist_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]


def test_fuse_multiscale_pair_equals_two_launches():
    """train1.py:410-424: target5 (64x64) and target0 (32x32) come from ONE launch (hp_fuse_multiscale_pair); the result is
    bit-identical to the two separate launches it replaced, on the driver's geometry and on one it does not cover."""
    fusion = importlib.import_module("domain-adaptative-hand-pose-estimation_b200.fusion")
    rs = np.random.RandomState(5101)
    for (s3, s2, hi, mid_out) in ((16, 32, 64, 32), (8, 16, 32, 16), (16, 32, 96, 48)):
        y3 = torch.from_numpy(rs.standard_normal((5, 21, s3, s3)).astype(np.float32)).cuda()
        y2 = torch.from_numpy(rs.standard_normal((5, 21, s2, s2)).astype(np.float32)).cuda()
        t5, t0 = hp.fuse_multiscale(y3, y2, hi, mid_out)
        w5 = fusion._fuse(y3, 0.5, y2, 1.0, None, 0.0, hi)
        w0 = fusion._fuse(y3, 1.0, None, 0.0, None, 0.0, mid_out)
        assert torch.equal(t5, w5) and torch.equal(t0, w0), (s3, s2, hi, mid_out)
        r5, r0 = O.fuse_multiscale(y3.cpu(), y2.cpu(), hi, mid_out)
        # exact x2 / x4 scales: the tap weights are dyadic, only the association order differs from ATen's (1e-6).  The x3 / x6
        # geometry has source indices up to 15.9 computed in fp32 (one ulp = 9.5e-7 of a weight times |v1 - v0| <= 6): 1e-5.
        atol = 1e-6 if hi % s3 == 0 and (hi // s3) in (2, 4) else 1e-5
        np.testing.assert_allclose(t5.cpu().numpy(), r5.numpy(), rtol=1e-5, atol=atol)
        np.testing.assert_allclose(t0.cpu().numpy(), r0.numpy(), rtol=1e-5, atol=atol)
