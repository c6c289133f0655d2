"""GPU: the small boundary operators of csrc/hp_extras.cu.

* float64 heatmaps through get_max_preds / accuracy (utils/keypoint_detection.py:12-21 takes any ndarray dtype): checked
  against the oracle, including a map whose two largest values differ only below float32 resolution (a cast would tie them
  and move the argmax);
* the OPT-IN quarter-pixel refinement (row a13: named by the north star, absent from the reference): hand-derived known
  answers + a numpy statement of the standard rule; default off == the reference bit for bit;
* group_accuracy on the device (keypoint_dataset.py:58-71): bit-equal to the reference's sum(...)/len(...)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import hp_oracle as O

pytestmark = pytest.mark.gpu
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")


def test_float64_heatmaps_decode_and_accuracy_match_the_oracle():
    rs = np.random.RandomState(3101)
    B, K, H, W = 5, 21, 64, 64
    out = rs.standard_normal((B, K, H, W))                       # float64
    tgt = rs.standard_normal((B, K, H, W))
    # two maxima that only float64 can tell apart: the LATER one is larger by 1e-12 -> argmax must be the later index
    out[0, 0].flat[100] = 10.0
    out[0, 0].flat[2000] = 10.0 + 1e-12
    assert np.float32(out[0, 0].flat[100]) == np.float32(out[0, 0].flat[2000])
    out[1, 1] = -np.abs(out[1, 1])                                # all <= 0 -> (0, 0)
    out[2, 2].flat[77] = np.nan                                   # NaN wins, masked to (0, 0)
    out[2, 3].flat[5] = np.nan; out[2, 3].flat[3] = np.nan        # the FIRST NaN
    want_xy, want_max = O.get_max_preds(out)
    got_xy, got_max = hp.get_max_preds(out)
    assert got_max.dtype == np.float64 and got_xy.dtype == np.float32
    assert np.array_equal(got_xy, want_xy)
    assert np.array_equal(got_max, want_max, equal_nan=True)
    assert tuple(got_xy[0, 0]) == (2000 % W, 2000 // W)
    acc, avg, cnt, pred = hp.accuracy(out, tgt)
    w_acc, w_avg, w_cnt, w_pred = O.accuracy(out, tgt)
    assert np.array_equal(acc, w_acc) and avg == w_avg and cnt == w_cnt and np.array_equal(pred, w_pred)
    # CUDA float64 tensors take the same path without a host hop
    t_xy, t_max = hp.get_max_preds(torch.from_numpy(out).cuda())
    assert t_max.dtype == torch.float64 and np.array_equal(t_xy.cpu().numpy(), want_xy)


def _refine_numpy(heat, preds):
    out = preds.copy()
    B, K, H, W = heat.shape
    for b in range(B):
        for k in range(K):
            px, py = int(np.floor(preds[b, k, 0] + 0.5)), int(np.floor(preds[b, k, 1] + 0.5))
            if 1 < px < W - 1 and 1 < py < H - 1:
                hm = heat[b, k]
                d = np.array([hm[py, px + 1] - hm[py, px - 1], hm[py + 1, px] - hm[py - 1, px]], dtype=np.float32)
                out[b, k] += np.sign(d) * np.float32(0.25)
    return out


def test_quarter_pixel_refinement_is_opt_in_and_follows_the_standard_rule():
    H = W = 16
    hm = np.zeros((1, 6, H, W), np.float32)
    hm[0, 0, 5, 7] = 1.0; hm[0, 0, 5, 8] = 0.5; hm[0, 0, 4, 7] = 0.25            # right / up neighbours higher
    hm[0, 1, 5, 7] = 1.0; hm[0, 1, 5, 6] = 0.5; hm[0, 1, 6, 7] = 0.25            # left / down
    hm[0, 2, 5, 7] = 1.0                                                        # symmetric: no move
    hm[0, 3, 1, 7] = 1.0; hm[0, 3, 1, 8] = 0.5                                  # y == 1: not interior
    hm[0, 4, 5, W - 1] = 1.0                                                    # on the border
    hm[0, 5] = -1.0                                                             # masked map stays (0, 0)
    want = np.array([[7.25, 4.75], [6.75, 5.25], [7.0, 5.0], [7.0, 1.0], [W - 1.0, 5.0], [0.0, 0.0]], np.float32)
    plain, _ = hp.get_max_preds(hm)
    ref_plain, _ = O.get_max_preds(hm)
    assert np.array_equal(plain, ref_plain)                                     # default: the reference, bit for bit
    got, _ = hp.get_max_preds(hm, refine="quarter")
    assert np.array_equal(got[0], want)
    d = hp.synth.make_host_batch(3102, 6)
    base, _ = O.get_max_preds(d["pred"])
    got, _ = hp.get_max_preds(d["pred"], refine="quarter")
    assert np.array_equal(got, _refine_numpy(d["pred"], base))
    t, _ = hp.decode(torch.from_numpy(d["pred"]).cuda(), refine="quarter")
    assert np.array_equal(t.cpu().numpy(), got)
    with pytest.raises(ValueError):
        hp.get_max_preds(hm, refine="half")


def test_group_accuracy_on_device_is_bit_equal_to_the_reference_rule():
    rs = np.random.RandomState(3103)
    acc = rs.uniform(size=21)
    acc[3] = -1.0                                               # a joint without valid targets (accuracy() gives -1)
    groups = {"MCP": (1, 5, 9, 13, 17), "PIP": (2, 6, 10, 14, 18), "DIP": (3, 7, 11, 15, 19),
              "fingertip": (4, 8, 12, 16, 20), "all": tuple(range(21))}        # keypoint_dataset.py:109-120
    want = {n: sum([acc[i] for i in idx]) / len(idx) for n, idx in groups.items()}   # keypoint_dataset.py:68-70
    got = hp.group_accuracy(torch.from_numpy(acc).cuda(), groups)
    assert got == want
    got2 = hp.group_accuracy(acc, groups)
    assert got2 == want
    # fed straight from the device-side PCK
    d = hp.synth.make_host_batch(3104, 8)
    tgt, _ = O.generate_target_batch(d["joints"], d["vis"], (64, 64), 2, (256, 256))
    acc_vec, _, _ = hp.pck(torch.from_numpy(d["pred"]).cuda(), torch.from_numpy(tgt).cuda())
    w_acc, _, _, _ = O.accuracy(d["pred"], tgt)
    got3 = hp.group_accuracy(acc_vec[:21], groups)
    assert got3 == {n: sum([w_acc[i] for i in idx]) / len(idx) for n, idx in groups.items()}
