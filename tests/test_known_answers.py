"""Hand-derived known answers for the oracle (the reference ships no tests, so besides the bit-equality pin
against the real reference the oracle is held to answers worked out by hand from the reference's source).  CPU only."""
import numpy as np
import torch

from oracle import hp_oracle as O


def _maps(h=64, w=64, n=1, k=1):
    return np.zeros((n, k, h, w), dtype=np.float32)


def test_decode_single_peak_ties_nonpositive_and_nan():
    m = _maps(k=5)
    m[0, 0, 17, 42] = 0.7                       # single peak -> (x, y) = (42, 17)
    m[0, 1, 5, 9] = m[0, 1, 5, 3] = 2.0         # duplicate maximum -> lowest flat index (row 5, col 3)
    m[0, 2] = -1.0                              # all <= 0 -> coordinates masked to (0, 0), max reported
    m[0, 3, 60, 1] = 1e-30                      # tiny but positive -> kept
    m[0, 4, 2, 2] = np.nan                      # NaN is numpy's argmax; max is NaN -> "max > 0" false -> (0, 0)
    m[0, 4, 30, 30] = 5.0
    xy, mx = O.get_max_preds(m)
    assert xy.dtype == np.float32 and xy.shape == (1, 5, 2) and mx.shape == (1, 5, 1)
    assert xy[0].tolist() == [[42.0, 17.0], [3.0, 5.0], [0.0, 0.0], [1.0, 60.0], [0.0, 0.0]]
    assert mx[0, 0, 0] == np.float32(0.7) and mx[0, 1, 0] == 2.0 and mx[0, 2, 0] == -1.0 and np.isnan(mx[0, 4, 0])


def test_pck_threshold_is_10_vs_11_squared_pixels_at_64():
    # norm = 6.4 per axis; hit iff sqrt(dx^2 + dy^2) / 6.4 < 0.5  <=>  dx^2 + dy^2 < 10.24
    tgt = _maps(n=3)
    out = _maps(n=3)
    for s, (dx, dy) in enumerate([(3, 1), (3, 2), (0, 0)]):       # 10 -> hit, 13 -> miss, 0 -> hit
        tgt[s, 0, 20, 20] = 1.0
        out[s, 0, 20 + dy, 20 + dx] = 1.0
    acc, avg, cnt, pred = O.accuracy(out, tgt)
    assert cnt == 1 and acc[0] == 2.0 / 3.0 and avg == 2.0 / 3.0
    hits, valid = O.pck_counts(pred, O.get_max_preds(tgt)[0], 64, 64)
    assert hits.tolist() == [2] and valid.tolist() == [3]
    out2 = _maps()
    out2[0, 0, 21, 23] = 1.0                                       # dx^2 + dy^2 = 9 + 1 = 10 -> hit ...
    out3 = _maps()
    out3[0, 0, 21, 23 + 0] = 0.0
    out3[0, 0, 23, 22] = 1.0                                       # ... 4 + 9 = 13 -> miss; (1, 3): 1 + 9 = 10 -> hit
    t1 = _maps()
    t1[0, 0, 20, 20] = 1.0
    assert O.accuracy(out2, t1)[0][0] == 1.0 and O.accuracy(out3, t1)[0][0] == 0.0


def test_pck_ignores_targets_at_or_below_one():
    tgt = _maps(k=3)
    out = _maps(k=3)
    tgt[0, 0, 1, 30] = 1.0          # y == 1 -> not "> 1": ignored
    tgt[0, 1, 30, 1] = 1.0          # x == 1: ignored
    tgt[0, 2, 2, 2] = 1.0           # both 2: valid
    out[0, :, 2, 2] = 1.0
    acc, avg, cnt, _ = O.accuracy(out, tgt)
    assert acc.tolist() == [-1.0, -1.0, 1.0] and cnt == 1 and avg == 1.0
    acc0, avg0, cnt0, _ = O.accuracy(out[:, :2], tgt[:, :2])      # nothing valid at all
    assert acc0.tolist() == [-1.0, -1.0] and cnt0 == 0 and avg0 == 0


def test_generated_target_centre_truncation_edges_and_weights():
    joints = np.array([[128.0, 128.0],      # mu = int(32 + 0.5) = 32
                       [-5.9, -5.9],        # -1.475 + 0.5 = -0.975 -> int() truncates toward zero -> 0 (in bounds)
                       [-6.1, 2.0],         # -1.525 + 0.5 = -1.025 -> -1 -> out of bounds: weight 0, map stays 0
                       [255.9, 255.9],      # 63.975 + 0.5 -> 64 -> out of bounds
                       [0.0, 252.0],        # corner-ish: mu = (0, 63): truncated patch
                       [128.0, 128.0]],     # in bounds but invisible: weight stays 0.4, nothing pasted (not > 0.5)
                      dtype=np.float64)
    vis = np.array([[1.0], [1.0], [1.0], [1.0], [1.0], [0.4]], dtype=np.float32)
    t, w = O.generate_target(joints, vis, (64, 64), 2, (256, 256))
    assert t.dtype == np.float32 and t.shape == (6, 64, 64) and w.shape == (6, 1)
    assert w[:, 0].tolist() == [1.0, 1.0, 0.0, 0.0, 1.0, np.float32(0.4)]
    assert t[0, 32, 32] == 1.0 and t[0].argmax() == 32 * 64 + 32
    assert np.count_nonzero(t[0]) == 13 * 13                       # sigma 2 -> tmp 6 -> 13x13 patch, all values > 0
    assert t[0, 32, 38] == np.exp(np.float32(-36.0 / 8.0)) and t[0, 32, 39] == 0.0
    assert t[1, 0, 0] == 1.0 and np.count_nonzero(t[1]) == 7 * 7   # centre (0, 0): only the lower-right 7x7 survives
    assert not t[2].any() and not t[3].any() and not t[5].any()
    assert t[4, 63, 0] == 1.0 and np.count_nonzero(t[4]) == 7 * 7


def test_loss_reductions_shapes_and_closed_forms():
    B, K, H, W = 2, 3, 8, 8
    pred = torch.zeros(B, K, H, W)
    tgt = torch.zeros(B, K, H, W)
    tgt[:, :, 0, 0] = 1.0
    w = torch.tensor([[1.0, 0.0, 2.0], [1.0, 1.0, 1.0]]).view(B, K, 1)
    # MSE: 0.5 * w * (p - t)^2, one pixel differs by 1 -> per map 0.5 * w / 64; 'mean' keeps zero-weight maps in the denominator
    none = O.joints_mse_loss(pred, tgt, w, "none")
    assert tuple(none.shape) == (B, K)
    assert torch.allclose(none, 0.5 * w.view(B, K) / 64)
    assert torch.isclose(O.joints_mse_loss(pred, tgt, w, "mean"), (0.5 * w.view(B, K) / 64).mean())
    # KL with a uniform prediction (log_softmax = -ln 64) and a one-hot target (eps = 0): sum q ln q - sum q logp = 0 + ln 64
    kl_none = O.joints_kl_loss(pred, tgt, w, "none", 0.0)
    assert tuple(kl_none.shape) == (B,)                            # the reference returns [B] (mean over K), not [B, K]
    want = np.log(64.0) * w.view(B, K).mean(dim=1)
    assert torch.allclose(kl_none, want.float(), rtol=1e-6)
    assert torch.isclose(O.joints_kl_loss(pred, tgt, w, "mean", 0.0), torch.tensor(np.log(64.0) * 6.0 / 6.0).float(), rtol=1e-6)   # weights sum to 6 over 6 maps
    # eps = 0 and an all-zero target: 0/0 -> NaN, even where the weight is 0
    tgt[0, 1] = 0.0
    assert torch.isnan(O.joints_kl_loss(pred, tgt, w, "mean", 0.0))
    assert not torch.isnan(O.joints_kl_loss(pred, tgt, w, "mean", 1e-7))


def test_bilinear_fusion_weights_are_the_dyadic_taps():
    # x2 upsample, align_corners=False: out[0] = in[0]; out[1] = .75 in[0] + .25 in[1]; out[2] = .25 in[0] + .75 in[1]; ...
    lo = torch.zeros(1, 1, 2, 2)
    lo[0, 0, 0, 0] = 4.0
    t5, t0 = O.fuse_multiscale(torch.zeros(1, 1, 1, 1), lo, 4, 2)    # target5 = 0.5 * up4(1x1 zeros) + up4(lo)
    row0 = t5[0, 0, 0].tolist()
    assert row0 == [4.0, 3.0, 1.0, 0.0]
    assert t5[0, 0, 1].tolist() == [3.0, 2.25, 0.75, 0.0]
    assert tuple(t0.shape) == (1, 1, 2, 2) and not t0.any()
