"""CPU tests: the front-end / render oracle against OpenCV itself (the library the reference calls), and the host-side
index tables of libespnet_b200.so against the oracle."""
import numpy as np
import pytest

from oracle import frontend_oracle as F
from oracle import wsi_oracle as W

cv2 = pytest.importorskip("cv2")


def test_add_weighted_matches_cv2_exhaustively():
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    assert np.array_equal(F.add_weighted_04_06(a, b), cv2.addWeighted(a, 0.4, b, 0.6, 0))


@pytest.mark.parametrize("src,dst", [(517, 200), (333, 700), (512, 1024), (100, 100), (7, 3), (3, 11)])
def test_nearest_indices_match_cv2(src, dst):
    ramp = np.arange(src, dtype=np.float32).reshape(1, src)
    got = cv2.resize(ramp, (dst, 1), interpolation=cv2.INTER_NEAREST).reshape(-1).astype(np.int64)
    assert np.array_equal(W.nearest_index(dst, src), got)


@pytest.mark.parametrize("hw,WH", [((333, 517), (200, 700)), ((333, 517), (1024, 512)), ((96, 80), (96, 80)), ((50, 61), (300, 300))])
def test_bilinear_restatement_vs_opencv_generic_path(hw, WH):
    """Within 2 float32 ulps of OpenCV's generic C++ path (bit-equal on most shapes; OpenCV's own IPP path is 1e-4 away)."""
    rng = np.random.default_rng(hw[0] + WH[0])
    img = rng.normal(size=hw + (3,)).astype(np.float32)
    prev = cv2.useOptimized()
    cv2.setUseOptimized(False)         # OpenCV's own generic C++ path; its IPP path differs from it by ~1e-4
    try:
        ref = cv2.resize(img, WH)
    finally:
        cv2.setUseOptimized(prev)
    assert np.allclose(F.resize_linear_f32(img, *WH), ref, rtol=3e-7, atol=3e-7)
    # the accelerated path stays within 2e-4 relative of the restatement
    fast = cv2.resize(img, WH)
    assert np.abs(fast - ref).max() <= 2e-4 * np.abs(ref).max()


def test_preprocess_and_colorize_follow_the_reference_lines():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (70, 90, 3), dtype=np.uint8)
    mean, std = (204.60071, 170.19359, 199.57469), (20.61257, 42.92207, 28.401505)
    # the reference's own statements, executed with cv2 (generic path)
    x = img.astype(np.float32)
    for j in range(3):
        x[:, :, j] -= mean[j]
    for j in range(3):
        x[:, :, j] /= std[j]
    prev = cv2.useOptimized()
    cv2.setUseOptimized(False)
    try:
        x = cv2.resize(x, (128, 64))
    finally:
        cv2.setUseOptimized(prev)
    x /= 255
    assert np.allclose(F.preprocess_resize(img, mean, std, 128, 64), x.transpose((2, 0, 1)), rtol=3e-7, atol=1e-9)
    lab = rng.integers(0, 27, (40, 50)).astype(np.uint8)
    col = np.zeros(lab.shape + (3,), np.uint8)
    for idx in range(len(F.PALLETE)):
        r, g, b = F.PALLETE[idx]
        col[lab == idx] = [b, g, r]
    assert np.array_equal(F.colorize(lab), col)
    assert np.array_equal(F.class_counts(lab, 5), [np.count_nonzero(lab == k) for k in range(5)])


def test_host_luts_of_the_library_match_the_oracle():
    from glomeruli_segmentation_b200 import frontend
    for src, dst in [(517, 200), (333, 700), (512, 1024), (64, 64), (9, 2)]:
        idx, wgt = frontend.bilinear_lut(src, dst)
        sx, fx = F.bilinear_coords(dst, src)
        assert np.array_equal(idx, sx) and np.array_equal(wgt, fx)
        assert np.array_equal(frontend.nearest_lut(src, dst), W.nearest_index(dst, src))
    assert frontend.PALLETE == F.PALLETE
