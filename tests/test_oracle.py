"""CPU tests: the oracle (oracle/) against the golden fixtures generated from the real reference
(tests/golden/make_golden.py) and, when /root/reference is mounted, against the live reference."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import espnet_oracle as O
from oracle import wsi_oracle as W

REF = "/root/reference"
have_ref = os.path.isdir(os.path.join(REF, "module/espnet/test"))


def test_p0_normalise_bit_exact(golden):
    for k in (1, 3):
        mean, std = O.FOLD_MEAN_STD[k]
        x = O.normalise_bgr_u8(golden["small_u8"], mean, std)
        assert x.dtype == np.float32
        assert np.array_equal(x, golden["small_x_fold%d" % k])


@pytest.mark.parametrize("fold", [1, 3])
def test_full_forward_matches_reference_golden(golden, fold_sd, fold):
    sd = fold_sd(fold)
    x = torch.from_numpy(golden["small_x_fold%d" % fold])
    y = O.espnet_forward(sd, x)
    ref = torch.from_numpy(golden["small_logits_fold%d" % fold])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 2e-5      # same ops, same library: only thread-order noise
    assert np.array_equal(O.argmax_mask(y), golden["small_mask_fold%d" % fold])


def test_mid_crop_taps_and_encoder(golden, fold_sd):
    sd = fold_sd(1)
    mean, std = O.FOLD_MEAN_STD[1]
    x = torch.from_numpy(O.normalise_bgr_u8(golden["mid_u8"], mean, std))
    taps = {}
    y = O.espnet_forward(sd, x, taps)
    assert (y - torch.from_numpy(golden["mid_logits_fold1"])).abs().max().item() <= 5e-5
    names = {"encoder.level1": "level1", "encoder.b1": "b1", "encoder.level2_0": "level2_0",
             "encoder.level2.0": "level2.0", "encoder.level2.1": "level2.1", "encoder.b2": "b2",
             "encoder.level3_0": "level3_0", "encoder.level3.0": "level3.0", "encoder.level3.7": "level3.7",
             "encoder.b3": "b3", "encoder.classifier": "encoder.classifier", "up_l3": "up_l3",
             "level3_C": "level3_C", "combine_l2_l3": "combine_l2_l3", "up_l2": "up_l2", "conv": "conv"}
    for gname, oname in names.items():
        g = golden["mid_tap:" + gname]
        tol = 5e-5 if g.dtype == np.float32 else 3e-2
        d = np.abs(taps[oname].numpy() - g.astype(np.float32)).max()
        assert d <= tol, (gname, d)
    e = O.espnet_encoder_forward(O.encoder_state_dict(sd), x)
    assert (e - torch.from_numpy(golden["mid_enc_logits_fold1"])).abs().max().item() <= 5e-5
    m = O.argmax_mask(O.upsample8_bilinear(e))
    assert (m == golden["mid_enc_mask_fold1"]).mean() >= 0.9999


def test_ensemble_matches_golden(golden, fold_sd):
    sds = [fold_sd(k) for k in range(1, 6)]
    mask, prob = O.ensemble_mask(sds, golden["ens_u8"], range(1, 6))
    assert np.abs(prob.numpy() - golden["ens_prob"]).max() <= 1e-5
    assert (mask == golden["ens_mask"]).mean() >= 0.9999


def test_random_state_dict_has_reference_key_set(fold_sd):
    sd = O.random_state_dict(5, 2, 8, seed=3)
    ref = fold_sd(1)
    assert set(sd) == set(ref)
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k


@pytest.mark.skipif(not have_ref, reason="/root/reference not mounted")
def test_oracle_against_live_reference(fold_sd):
    sys.path.insert(0, os.path.join(REF, "module/espnet/test"))
    import Model as RefModel
    sd = O.random_state_dict(5, 2, 3, seed=9)
    m = RefModel.ESPNet(5, 2, 3)
    m.load_state_dict(sd, strict=True)
    m.eval()
    x = torch.from_numpy(O.normalise_bgr_u8(O.synth_crops("D1", 2, 40, 72, seed=4), *O.FOLD_MEAN_STD[2]))
    with torch.no_grad():
        ref = m(x)
    assert (O.espnet_forward(sd, x) - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------------ WSI index work
def test_tiler_known_answers():
    # SURVEY.md 8(a) T1 [computed]: 40000x30000, STD=512, mpp=1, ds=1
    o, nx, ny, wx, wy, sx, sy = W.tile_grid(40000, 30000, 512, 1.0, 1.0, 0.1, 1.0)
    assert (nx, ny, wx, wy, sx, sy) == (87, 66, 512, 512, 460, 460) and len(o) == 5742
    assert tuple(o[1]) == (460, 0) and tuple(o[87]) == (0, 460) and tuple(o[-1]) == (460 * 86, 460 * 65)
    o, nx, ny, wx, wy, sx, sy = W.tile_grid(40000, 30000, 512, 1.0, 1.0, 0.5, 1.0)
    assert (nx, ny, sx) == (157, 118, 256) and len(o) == 18526
    # float mpp: python float semantics of ceil / int
    o, nx, ny, wx, wy, sx, sy = W.tile_grid(53248, 23040, 2000, 0.2277, 0.2277, 0.1, 8.0)
    assert nx == int(math.ceil(53248 / (2000 / 0.2277) / 0.9)) and wx == int(math.ceil(2000 / 0.2277 / 8.0))
    assert sx == int((2000 / 0.2277) * 0.9)


def test_stitch_windows_known_answers_and_quirk():
    w = W.stitch_windows(40000, 30000, 2400)
    assert len(w) == 17 * 13
    assert w[0] == (0, 0, 2400, 2400) and w[12] == (0, 28800, 2400, 30000) and w[-1] == (38400, 28800, 40000, 30000)
    # the reference compares ymax against the slide WIDTH: a tall narrow slide loses its lower windows
    w = W.stitch_windows(1000, 3000, 400)
    assert max(b[3] for b in w) <= 1000
    # W % ws == 0 emits a zero-width window (reference would crash there)
    assert any(b[0] == b[2] for b in W.stitch_windows(800, 500, 400))


def test_nearest_index_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (sw, sh, dw, dh) in [(2400, 2400, 300, 300), (1600, 1200, 200, 150), (37, 53, 4, 6), (1001, 999, 125, 124),
                             (64, 64, 512, 512), (100, 80, 33, 77)]:
        a = rng.integers(0, 5, (sh, sw)).astype(np.uint8)
        assert np.array_equal(W.resize_nearest(a, dw, dh), cv2.resize(a, (dw, dh), interpolation=cv2.INTER_NEAREST))


def test_overlay_and_stitch_small():
    rng = np.random.default_rng(1)
    Wd, Hd, ws = 1000, 760, 240
    boxes, masks = [], []
    for _ in range(40):
        x0, y0 = int(rng.integers(-30, Wd - 20)), int(rng.integers(-30, Hd - 20))
        w, h = int(rng.integers(8, 200)), int(rng.integers(8, 200))
        boxes.append([float(x0), float(y0), float(x0 + w), float(y0 + h), 0.9])
        masks.append(rng.integers(0, 5, (h, w)))
    level0, ds8 = W.stitch_slide(boxes, masks, Wd, Hd, ws)
    # independent statement: per-pixel max over covering boxes, clipped to the slide
    exp = np.zeros((Hd, Wd), int)
    for b, m in zip(boxes, masks):
        x0, y0, x1, y1 = map(int, b[:4])
        cx0, cy0, cx1, cy1 = max(x0, 0), max(y0, 0), min(x1, Wd), min(y1, Hd)
        if cx1 > cx0 and cy1 > cy0:
            exp[cy0:cy1, cx0:cx1] = np.maximum(exp[cy0:cy1, cx0:cx1], m[cy0 - y0:cy1 - y0, cx0 - x0:cx1 - x0])
    assert np.array_equal(level0, exp)      # W > H here, so the ymax>width quirk does not bite
    assert ds8.shape == (95, 125)
    assert np.array_equal(ds8[:90, :120], exp[:720:8, :960:8])


def test_iou_against_reference_golden():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "iou_golden.npz"))
    hist = sum(W.fast_hist(z["gt"][i], z["pred"][i], 5) for i in range(3))
    assert np.array_equal(hist, z["hist"])
    overall, per_acc, per_iou, miou = W.metric_right(hist)
    assert np.allclose(per_iou, z["per_iou"]) and np.isclose(miou, z["miou"]) and np.isclose(overall, z["overall"])
