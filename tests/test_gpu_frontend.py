"""GPU parity tests of the crop front-end / render kernels against the CPU oracle, through the C ABI.
u8 / index work: bit-exact.  Bilinear float resize: the kernel repeats the oracle's operations with non-fused fp32 ops, so it
is held to bit equality with the oracle (which is within 2 ulp of OpenCV's generic path)."""
import numpy as np
import pytest
import torch

from glomeruli_segmentation_b200 import ESPNet, FOLD_MEAN_STD, frontend, wsi
from oracle import espnet_oracle as O
from oracle import frontend_oracle as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("hw,WH", [((333, 517), (200, 700)), ((300, 421), (1024, 512)), ((64, 96), (96, 64)), ((17, 9), (8, 8))])
def test_preprocess_resize_bit_equal_to_oracle(hw, WH):
    rng = np.random.default_rng(hw[0])
    crops = rng.integers(0, 256, (2,) + hw + (3,), dtype=np.uint8)
    mean, std = FOLD_MEAN_STD[2]
    got = frontend.preprocess_resize(torch.from_numpy(crops).to(DEV), mean, std, WH[0], WH[1]).cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], F.preprocess_resize(crops[b], mean, std, WH[0], WH[1]))


@pytest.mark.parametrize("src,dst", [((512, 1024), (333, 517)), ((64, 96), (200, 100)), ((8, 8), (8, 8))])
def test_mask_nearest_resize_bit_exact(src, dst):
    rng = np.random.default_rng(src[0] + dst[0])
    m = rng.integers(0, 5, (3,) + src, dtype=np.uint8)
    got = frontend.resize_mask_nearest(torch.from_numpy(m).to(DEV), dst[0], dst[1]).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], F.resize_nearest(m[b], dst[1], dst[0]))


def test_palette_overlay_and_counts_bit_exact():
    rng = np.random.default_rng(9)
    lab = rng.integers(0, 30, (2, 130, 170)).astype(np.uint8)      # includes labels past the palette (stay black)
    img = rng.integers(0, 256, (2, 130, 170, 3), dtype=np.uint8)
    color, over = frontend.colorize_overlay(torch.from_numpy(lab).to(DEV), torch.from_numpy(img).to(DEV))
    ref_c = F.colorize(lab)
    assert np.array_equal(color.cpu().numpy(), ref_c)
    assert np.array_equal(over.cpu().numpy(), F.add_weighted_04_06(img, ref_c))
    cnt = frontend.class_pixel_counts(torch.from_numpy(lab).to(DEV), 5).cpu().numpy()
    for b in range(2):
        assert np.array_equal(cnt[b], F.class_counts(lab[b], 5))


@pytest.mark.parametrize("sw,sh,ws", [(1000, 760, 240), (500, 1300, 160), (333, 257, 80)])
def test_render_ds8_bit_exact(sw, sh, ws):
    rng = np.random.default_rng(sw)
    slide = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    level0 = np.kron(rng.integers(0, 5, (sh // 16 + 1, sw // 16 + 1)).astype(np.uint8), np.ones((16, 16), np.uint8))[:sh, :sw]
    # rows the reference's window loop never writes stay zero in the level-0 map as well (wsi.stitch_* honour y_limit)
    level0 = np.ascontiguousarray(level0)
    level0[wsi.stitch_y_limit(sw, sh, ws):] = 0
    got = frontend.render_slide_ds8(torch.from_numpy(slide).to(DEV), torch.from_numpy(level0).to(DEV), ws).cpu().numpy()
    assert np.array_equal(got, F.render_ds8(slide, level0, ws))


def test_non_native_crop_end_to_end(fold_sd):
    """The reference's per-crop loop for a crop that is not the network size (VisualizeResults_iou.py:103-129):
    normalise + bilinear resize -> forward -> arg-max -> nearest resize back, GPU vs oracle composition."""
    sd = fold_sd(1)
    mean, std = FOLD_MEAN_STD[1]
    crop = O.synth_crops("D2", 1, 203, 317, seed=4, sigma=3.0)[0]
    W, H = 256, 128
    x_ref = F.preprocess_resize(crop, mean, std, W, H)[None]
    ref_logits = O.espnet_forward(sd, torch.from_numpy(x_ref))
    ref_mask = F.resize_nearest(O.argmax_mask(ref_logits)[0], crop.shape[1], crop.shape[0])
    m = ESPNet(5, 2, 8)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = frontend.preprocess_resize(torch.from_numpy(crop[None]).to(DEV), mean, std, W, H)
    assert np.array_equal(x.cpu().numpy(), x_ref)
    y = m(x)
    assert (y.cpu() - ref_logits).abs().max().item() <= 1e-3
    mask = frontend.resize_mask_nearest(y.max(1)[1].byte(), crop.shape[0], crop.shape[1])[0].cpu().numpy()
    assert (mask == ref_mask).mean() >= 0.9999
