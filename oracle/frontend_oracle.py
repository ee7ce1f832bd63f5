"""CPU oracle for the crop front-end and the render steps next to the ESPNet forward -- TEST INFRASTRUCTURE ONLY
(only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it).

Restates, in numpy, the OpenCV calls the reference makes at
  module/espnet/test/VisualizeResults_iou.py:107-119  (normalise + cv2.resize INTER_LINEAR + /255 + HWC->CHW)
  module/espnet/test/VisualizeResults_iou.py:129      (cv2.resize INTER_NEAREST of the class map)
  module/espnet/test/VisualizeResults_iou.py:139-155  (palette colour map, cv2.addWeighted, class pixel counts)
  module/espnet/test/eval_wsi_segmentation.py:215-241, 359-394 (the /8 rendered slide)
The arithmetic lives in OpenCV (third party, absent from /root/reference; reference pin: `opencv-python` unversioned in
docker/requirements.txt; this image: cv2 4.13).  Pinned in tests/test_oracle_frontend.py against cv2 itself:
addWeighted exhaustively over all u8 pairs (bit-exact), INTER_NEAREST indices (bit-exact), and INTER_LINEAR on float32
within 2 ulp of OpenCV's generic code path (cv2.setUseOptimized(False)); cv2's IPP path deviates from its own generic path by
~1e-4 relative, so that is the tolerance against a stock cv2 build.
"""
from __future__ import annotations

import numpy as np

from . import wsi_oracle as W

f32 = np.float32

# eval_wsi_segmentation.py:23-47 (== VisualizeResults_iou.py:22-46), rows are [r, g, b]
PALLETE = [[0, 0, 0], [255, 0, 0], [0, 184, 0], [255, 255, 0], [0, 0, 255], [128, 64, 128], [244, 35, 232], [70, 70, 70],
           [102, 102, 156], [190, 153, 153], [153, 153, 153], [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152],
           [70, 130, 180], [220, 20, 60], [255, 0, 0], [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230],
           [119, 11, 32], [0, 0, 0]]


def bilinear_coords(dst_n: int, src_n: int):
    """OpenCV resizeLinear (generic path): fx = (float)((dx + 0.5) * scale - 0.5), sx = cvFloor(fx), fx -= sx;
    sx < 0 -> (0, 0); sx >= src-1 -> (src-1, 0); scale = 1. / ((double)dst / src)."""
    scale = 1.0 / (np.float64(dst_n) / np.float64(src_n))
    fx = ((np.arange(dst_n, dtype=np.float64) + 0.5) * scale - 0.5).astype(f32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(f32)).astype(f32)
    lo = sx < 0
    fx[lo] = 0
    sx[lo] = 0
    hi = sx >= src_n - 1
    fx[hi] = 0
    sx[hi] = src_n - 1
    return sx, fx


def resize_linear_f32(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """cv2.resize(img_float32_HWC, (width, height)) (INTER_LINEAR default): horizontal pass, then vertical pass, every
    product and sum rounded to float32 (HResizeLinear / VResizeLinear of the generic path)."""
    h, w = img.shape[:2]
    xs, fx = bilinear_coords(width, w)
    ys, fy = bilinear_coords(height, h)
    x1 = np.minimum(xs + 1, w - 1)
    y1 = np.minimum(ys + 1, h - 1)
    ax = (f32(1) - fx).astype(f32)
    ay = (f32(1) - fy).astype(f32)
    rows = (img[:, xs, :] * ax[None, :, None]).astype(f32) + (img[:, x1, :] * fx[None, :, None]).astype(f32)
    rows = rows.astype(f32)
    out = (rows[ys] * ay[:, None, None]).astype(f32) + (rows[y1] * fy[:, None, None]).astype(f32)
    return out.astype(f32)


def preprocess_resize(img_u8: np.ndarray, mean, std, width: int, height: int) -> np.ndarray:
    """VisualizeResults_iou.py:107-119: BGR u8 [h,w,3] -> float32 [3,height,width]."""
    img = img_u8.astype(f32)                                   # :107
    for j in range(3):
        img[:, :, j] -= f32(mean[j])                           # :108-109
    for j in range(3):
        img[:, :, j] /= f32(std[j])                            # :110-111
    img = resize_linear_f32(img, width, height)                # :114
    img = (img / f32(255)).astype(f32)                         # :116
    return np.ascontiguousarray(img.transpose((2, 0, 1)))      # :117


def resize_nearest(a: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """VisualizeResults_iou.py:129."""
    return W.resize_nearest(a, dst_w, dst_h)


def colorize(label: np.ndarray) -> np.ndarray:
    """VisualizeResults_iou.py:139-143 / eval_wsi_segmentation.py:230-233: zeros, then map[label == idx] = [b, g, r]."""
    out = np.zeros(label.shape + (3,), np.uint8)
    for idx in range(len(PALLETE)):
        r, g, b = PALLETE[idx]
        out[label == idx] = [b, g, r]
    return out


def add_weighted_04_06(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """cv2.addWeighted(a, 0.4, b, 0.6, 0) on u8: saturate_cast<uchar>(cvRound(a*0.4f + b*0.6f)), float32, round-half-even."""
    v = (a.astype(f32) * f32(0.4)).astype(f32) + (b.astype(f32) * f32(0.6)).astype(f32)
    return np.clip(np.rint(v.astype(f32)), 0, 255).astype(np.uint8)


def render_ds8(slide: np.ndarray, level0: np.ndarray, ws: int) -> np.ndarray:
    """generate_pred_wsi's canvas (eval_wsi_segmentation.py:359-394 with generate_whole_img :215-241): per window the /8
    nearest resize of the slide pixels and of the label, palette, addWeighted, paste at [ymin//8:ymax//8, xmin//8:xmax//8]."""
    sh, sw = level0.shape
    whole = np.zeros((int(sh / W.MAGNIFICATION), int(sw / W.MAGNIFICATION), 3), np.uint8)            # :370
    for (xmin, ymin, xmax, ymax) in W.stitch_windows(sw, sh, ws):
        w, h = xmax - xmin, ymax - ymin
        if w == 0 or h == 0:
            continue
        dw, dh = int(w / W.MAGNIFICATION), int(h / W.MAGNIFICATION)                                  # :227-228
        region = slide[ymin:ymax, xmin:xmax]
        ys, xs = W.nearest_index(dh, h), W.nearest_index(dw, w)
        region_s = region[ys][:, xs]
        label_s = level0[ymin:ymax, xmin:xmax][ys][:, xs]
        over = add_weighted_04_06(region_s, colorize(label_s))                                      # :229-235
        whole[ymin // W.MAGNIFICATION:ymax // W.MAGNIFICATION, xmin // W.MAGNIFICATION:xmax // W.MAGNIFICATION] = over   # :236-240
    return whole


def class_counts(mask: np.ndarray, n: int = 5) -> np.ndarray:
    """VisualizeResults_iou.py:151-155."""
    return np.array([np.count_nonzero(mask == k) for k in range(n)], np.int64)
