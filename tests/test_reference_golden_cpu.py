"""The CPU oracle held against fixtures produced by RUNNING the reference's own scripts
(tests/golden/make_wsi_golden.py: detect_glomus_test.GlomusDetector.scan_region, eval_wsi_segmentation's
generate_pred_wsi / overlay / generate_whole_img, annotation_handler.check_overlap, make_seg_data.output_org_files,
VisualizeResults_iou.evaluateModel).  Integer work is compared bit for bit."""
import os

import numpy as np
import pytest
import torch

import wsi_cases as WC
from oracle import espnet_oracle as O
from oracle import frontend_oracle as F
from oracle import wsi_oracle as W

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def wz():
    return np.load(os.path.join(GOLD, "wsi_golden.npz"))


@pytest.fixture(scope="module")
def fz():
    return np.load(os.path.join(GOLD, "frontend_golden.npz"))


@pytest.mark.parametrize("k", range(len(WC.T1_CASES)))
def test_tiler_equals_reference_scan_region(wz, k):
    sw, sh, std, mx, my, ov, power, lds = WC.T1_CASES[k]
    calls = wz["t1_%d_calls" % k]                      # (x_start, y_start, level, window_x, window_y) per read_region
    level, ds = W.select_level(power, lds)
    assert ds == float(wz["t1_%d_downsample" % k]) and (calls[:, 2] == level).all()
    org, nx, ny, wx, wy, sx, sy = W.tile_grid(sw, sh, std, mx, my, ov, ds)
    assert len(org) == len(calls) == nx * ny
    assert np.array_equal(org, calls[:, :2])
    assert (calls[:, 3] == wx).all() and (calls[:, 4] == wy).all()


@pytest.mark.parametrize("name", list(WC.STITCH_CASES))
def test_stitcher_equals_reference_generate_pred_wsi(wz, name):
    c = WC.STITCH_CASES[name]
    boxes, masks = WC.stitch_inputs(name)
    ref_windows = [tuple(int(v) for v in r) for r in wz["s_%s_windows" % name]]
    windows = W.stitch_windows(c["sw"], c["sh"], c["ws"])
    if int(wz["s_%s_crashed" % name]):
        # W % ws == 0: the reference's loop reaches a zero-sized window and dies in cv2.resize; the
        # oracle lists the same windows up to and including that one, callers drop it
        n = len(ref_windows)
        assert windows[:n] == ref_windows and windows[n - 1][0] == windows[n - 1][2]
        return
    assert windows == ref_windows
    level0, ds8 = W.stitch_slide(boxes, masks, c["sw"], c["sh"], c["ws"])
    assert np.array_equal(level0, wz["s_%s_level0" % name])
    assert np.array_equal(ds8, wz["s_%s_ds8" % name])
    assert np.array_equal(F.render_ds8(WC.slide_rgb(name), level0, c["ws"]), wz["s_%s_render" % name])
    if name == "tall":           # `if ymax > slide_width: continue` (eval_wsi_segmentation.py:386, sic) really drops the lower windows
        assert max(w[3] for w in ref_windows) <= c["sw"] < c["sh"] and not level0[c["sw"]:].any()


def test_rect_overlap_equals_reference(wz):
    got = np.array([W.check_overlap(a, b) for a, b in WC.overlap_pairs()], np.float64)
    assert np.array_equal(got, wz["overlap_scores"])


def test_crop_regions_equal_reference_output_org_files(wz):
    boxes, _ = WC.stitch_inputs("wide")
    regions, names = W.crop_regions(boxes)
    assert np.array_equal(np.array([(x, y, 0, w, h) for x, y, w, h in regions], np.int64), wz["crop_calls"])
    assert [n + ".PNG" for n in names] == [str(s) for s in wz["crop_names"]]
    slide = WC.slide_rgb("wide")
    for (x, y, w, h), ref in zip(regions, wz["crop_bgr_sums"]):
        bgr = W.read_tile(slide, x, y, w, h)[..., ::-1]          # zero-padded read; PNG -> cv2.imread hands back BGR
        assert [int(bgr[..., ch].astype(np.int64).sum()) for ch in range(3)] == list(ref[:3])
        assert (int(bgr[0, 0, 0]), int(bgr[-1, -1, 2])) == (int(ref[3]), int(ref[4]))


@pytest.mark.parametrize("k", range(len(WC.FRONTEND_CASES)))
def test_frontend_equals_reference_evaluate_model(fz, fold_sd, k):
    """evaluateModel run for real (stock cv2 of this image, reference Model.py + shipped checkpoint): the oracle's
    normalise + bilinear resize within 1e-5 of the tensor's max-abs (stock cv2 takes its IPP/SIMD path: ~4e-6 seen; the
    oracle restates the generic path), the class maps, their nearest resize back and the pixel counts exactly."""
    ch, cw, in_w, in_h, fold, dist, seed = WC.FRONTEND_CASES[k]
    crops = O.synth_crops(dist, 2, ch, cw, seed=seed, sigma=3.0)
    mean, std = O.FOLD_MEAN_STD[fold]
    ref_in = fz["f_%d_net_in" % k]
    x = np.stack([F.preprocess_resize(c, mean, std, in_w, in_h) for c in crops])
    assert np.abs(x - ref_in).max() <= 1e-5 * np.abs(ref_in).max()
    if (ch, cw) == (in_h, in_w):
        assert np.array_equal(x, ref_in)                 # identity size: bit-equal (P0's three roundings)
    sd = fold_sd(fold)
    logits = O.espnet_forward(sd, torch.from_numpy(x))
    if "f_%d_logits" % k in fz.files:
        assert (logits - torch.from_numpy(fz["f_%d_logits" % k])).abs().max().item() <= 1e-3
    back = np.stack([F.resize_nearest(m, cw, ch) for m in O.argmax_mask(logits)])
    ref_masks = fz["f_%d_masks" % k]
    assert (back == ref_masks).mean() >= 0.9999
    assert np.array_equal(np.stack([F.class_counts(m) for m in ref_masks]), fz["f_%d_counts" % k])
