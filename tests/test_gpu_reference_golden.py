"""The CUDA path (through the C ABI) held against fixtures produced by RUNNING the reference's own scripts
(tests/golden/make_wsi_golden.py).  Tiling / stitching / render / nearest indices bit-exact; the bilinear crop
front-end within 1e-5 of the tensor's max-abs of what evaluateModel fed its network (stock cv2)."""
import os

import numpy as np
import pytest
import torch

import wsi_cases as WC
from glomeruli_segmentation_b200 import ESPNet, frontend, wsi
from oracle import espnet_oracle as O
from oracle import frontend_oracle as F
from oracle import wsi_oracle as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def wz():
    return np.load(os.path.join(GOLD, "wsi_golden.npz"))


@pytest.fixture(scope="module")
def fz():
    return np.load(os.path.join(GOLD, "frontend_golden.npz"))


@pytest.mark.parametrize("k", range(len(WC.T1_CASES)))
def test_tile_grid_equals_reference_scan_region(wz, k):
    sw, sh, std, mx, my, ov, power, lds = WC.T1_CASES[k]
    calls = wz["t1_%d_calls" % k]
    level, ds = wsi.select_level(power, lds)
    assert ds == float(wz["t1_%d_downsample" % k]) and (calls[:, 2] == level).all()
    g = wsi.tile_grid(sw, sh, std, mx, my, ov, ds)
    assert g.count == len(calls)
    assert np.array_equal(g.origins().astype(np.int64), calls[:, :2])
    assert (calls[:, 3] == g.win_x).all() and (calls[:, 4] == g.win_y).all()
    # band sharding enumerates the same tiles
    parts = [g.origins(*wsi.shard_rows(g.n_y, r, 3)) for r in range(3)]
    assert np.array_equal(np.concatenate(parts).astype(np.int64), calls[:, :2])


@pytest.mark.parametrize("name", [n for n in WC.STITCH_CASES if n != "divisible"])
def test_stitch_kernels_equal_reference_generate_pred_wsi(wz, name):
    c = WC.STITCH_CASES[name]
    boxes, masks = WC.stitch_inputs(name)
    got = torch.zeros((c["sh"], c["sw"]), dtype=torch.uint8, device=DEV)
    wsi.stitch_boxes(got, boxes, [torch.from_numpy(m).to(DEV) for m in masks], c["ws"])
    assert np.array_equal(got.cpu().numpy(), wz["s_%s_level0" % name])                         # T2 + T3
    assert np.array_equal(wsi.downsample8(got, c["ws"]).cpu().numpy(), wz["s_%s_ds8" % name])  # T4
    slide = torch.from_numpy(WC.slide_rgb(name)).to(DEV)
    assert np.array_equal(frontend.render_slide_ds8(slide, got, c["ws"]).cpu().numpy(), wz["s_%s_render" % name])   # f3


def test_box_crops_equal_reference_output_org_files(wz):
    """make_seg_data.output_org_files + VisualizeResults_iou.py:103-119 in one kernel, on the boxes and slide the reference
    script was run on: region arguments and names equal; the resized network input equals the oracle front-end applied to
    the cut crop (itself pinned to evaluateModel in tests/test_reference_golden_cpu.py) bit for bit."""
    boxes, _ = WC.stitch_inputs("wide")
    regions, names = frontend.crop_regions(boxes)
    assert np.array_equal(np.array([(x, y, 0, w, h) for x, y, w, h in regions], np.int64), wz["crop_calls"])
    assert [n + ".PNG" for n in names] == [str(s) for s in wz["crop_names"]]
    slide_bgr = np.ascontiguousarray(WC.slide_rgb("wide")[..., ::-1])
    mean, std = O.FOLD_MEAN_STD[1]
    sel = list(range(0, len(boxes), 3))
    got = frontend.preprocess_boxes(torch.from_numpy(slide_bgr).to(DEV), [boxes[i] for i in sel], mean, std, 256, 128).cpu().numpy()
    for j, i in enumerate(sel):
        x, y, w, h = regions[i]
        crop = W.read_tile(slide_bgr, x, y, w, h)
        ref = wz["crop_bgr_sums"][i]
        assert [int(crop[..., ch].astype(np.int64).sum()) for ch in range(3)] == list(ref[:3])
        exp = F.preprocess_resize(crop, mean, std, 256, 128)
        assert np.abs(got[j] - exp).max() <= 2e-7 * np.abs(exp).max()       # 2 ulp (FMA-free arithmetic on both sides)


@pytest.mark.parametrize("k", range(len(WC.FRONTEND_CASES)))
def test_crop_pipeline_equals_reference_evaluate_model(fz, fold_sd, k):
    """normalise + resize -> forward -> arg-max -> nearest resize back -> class counts, against what the reference's
    evaluateModel computed with stock cv2, its own Model.py and the shipped checkpoint."""
    ch, cw, in_w, in_h, fold, dist, seed = WC.FRONTEND_CASES[k]
    crops = torch.from_numpy(O.synth_crops(dist, 2, ch, cw, seed=seed, sigma=3.0)).to(DEV)
    mean, std = O.FOLD_MEAN_STD[fold]
    ref_in = fz["f_%d_net_in" % k]
    x = frontend.preprocess_resize(crops, mean, std, in_w, in_h)
    assert np.abs(x.cpu().numpy() - ref_in).max() <= 1e-5 * np.abs(ref_in).max()      # the bar held against stock (IPP) cv2
    m = ESPNet(5, 2, 8)
    m.load_state_dict(fold_sd(fold), strict=True)
    m = m.to(DEV).eval()
    logits = m(x)
    if "f_%d_logits" % k in fz.files:
        assert (logits.cpu() - torch.from_numpy(fz["f_%d_logits" % k])).abs().max().item() <= 1e-3
    small = logits.argmax(1).to(torch.uint8)
    back = frontend.resize_mask_nearest(small, ch, cw)
    ref_masks = fz["f_%d_masks" % k]
    assert (back.cpu().numpy() == ref_masks).mean() >= 0.9999
    counts = frontend.class_pixel_counts(torch.from_numpy(ref_masks).to(DEV))
    assert np.array_equal(counts.cpu().numpy(), fz["f_%d_counts" % k])
