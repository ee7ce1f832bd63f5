"""CPU oracle for the tile -> slide index arithmetic  --  TEST INFRASTRUCTURE ONLY.

numpy / pure-Python restatement of the integer work either side of the ESPNet forward:
the overlapping tiler (T1), the stitch windows (T2), the box paste + max merge (T3), the /8
nearest down-sample + paste (T4), the box -> crop extraction and the confusion-matrix IoU.
Each function follows the cited line range and reproduces Python `int()`, `//` and `math.ceil`
on floats exactly.

PINNED AGAINST THE REFERENCE ITSELF: tests/golden/make_wsi_golden.py imports the unmodified
reference scripts (tensorflow / openslide / labelme replaced by stub modules, an in-memory slide
behind openslide.open_slide) and runs the real GlomusDetector.scan_region,
Generate_Segmentation_Gt.generate_pred_wsi / overlay / generate_whole_img,
AnnotationHandler.check_overlap and make_seg_data's output_org_files on seeded slides
(tests/wsi_cases.py); tests/test_reference_golden_cpu.py holds this module equal to those
fixtures (tests/golden/wsi_golden.npz).  Only tests / smoke / bench's CPU legs may import this.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

MAGNIFICATION = 8  # eval_wsi_segmentation.py:22


# ------------------------------------------------------------------------------------------
# T1 -- overlapping sliding-window tiler
# ------------------------------------------------------------------------------------------
def tile_grid(slide_w: int, slide_h: int, std_size: float, mpp_x: float, mpp_y: float,
              overlap: float, downsample: float):
    """detect_glomus_test.py:286-304 (calc_window_size) + :264-275 (scan_region loop).

    Returns (origins[int64 N,2] as (x_start, y_start) in row-major `for j: for i:` order,
             n_x, n_y, win_x, win_y, stride_x, stride_y)."""
    win_x_org = float(std_size) / mpp_x                                   # :293
    win_y_org = float(std_size) / mpp_y                                   # :294
    n_x = int(math.ceil(slide_w / win_x_org / (1.0 - overlap)))           # :297
    n_y = int(math.ceil(slide_h / win_y_org / (1.0 - overlap)))           # :298
    win_x = int(math.ceil(win_x_org / downsample))                        # :301
    win_y = int(math.ceil(win_y_org / downsample))                        # :302
    stride_x = int(win_x_org * (1.0 - overlap))                           # :265
    stride_y = int(win_y_org * (1.0 - overlap))                           # :266
    origins = np.empty((n_x * n_y, 2), np.int64)
    k = 0
    for j in range(0, n_y):                                               # :268
        for i in range(0, n_x):                                           # :269
            origins[k, 0] = stride_x * i                                  # :270
            origins[k, 1] = stride_y * j                                  # :271
            k += 1
    return origins, n_x, n_y, win_x, win_y, stride_x, stride_y


def select_level(objective_power: float, level_downsamples: Sequence[float]):
    """detect_glomus_test.py:255-262: the first pyramid level whose magnification is <= 5x; when none
    is, the reference keeps its defaults `target_level = 3`, `slide_downsample = 8.0` (sic).
    Returns (target_level, slide_downsample)."""
    downsample, target_level = 8.0, 3                                     # :255-256
    for level, ds in enumerate(level_downsamples):                        # :257
        if objective_power / ds <= 5.0:                                   # :258
            target_level = level
            downsample = level_downsamples[level]
            break
    return target_level, downsample


def read_tile(slide: np.ndarray, x0: int, y0: int, win_x: int, win_y: int) -> np.ndarray:
    """`slide.read_region((x0,y0), level 0, (win_x, win_y))` on an in-memory [H,W,3] slide:
    out-of-bounds pixels are zero (openslide pads with transparent black; the alpha channel is
    dropped at detect_glomus_test.py:276)."""
    h, w = slide.shape[:2]
    out = np.zeros((win_y, win_x, slide.shape[2]), slide.dtype)
    cx0, cy0 = max(x0, 0), max(y0, 0)
    x1, y1 = min(x0 + win_x, w), min(y0 + win_y, h)
    if x1 > cx0 and y1 > cy0:
        out[cy0 - y0: y1 - y0, cx0 - x0: x1 - x0] = slide[cy0:y1, cx0:x1]
    return out


def crop_regions(boxes: Sequence[Sequence[float]]):
    """make_seg_data.py:347-361 (output_org_files): per detected box the level-0 `read_region` arguments
    (x, y, w, h) and the crop's file stem `xmin{}_ymin{}_xmax{}_ymax{}` in /8 coordinates (:359)."""
    regions, names = [], []
    for b in boxes:
        regions.append((b[0], b[1], b[2] - b[0], b[3] - b[1]))           # :358
        names.append("xmin{}_ymin{}_xmax{}_ymax{}".format(int(b[0] / MAGNIFICATION), int(b[1] / MAGNIFICATION),
                                                           int(b[2] / MAGNIFICATION), int(b[3] / MAGNIFICATION)))   # :359
    return regions, names


# ------------------------------------------------------------------------------------------
# T2 -- stitch windows
# ------------------------------------------------------------------------------------------
def stitch_windows(slide_w: int, slide_h: int, ws: int) -> List[Tuple[int, int, int, int]]:
    """eval_wsi_segmentation.py:180-195 (== :372-387), including the reference's
    `if ymax > slide_width: continue` (y compared against the WIDTH, sic).  Returns
    [(xmin, ymin, xmax, ymax)] in the reference's x-outer / y-inner order.  Zero-sized edge windows
    (W % ws == 0) are emitted too; the reference would crash in cv2.resize on them
     -- callers drop them (the SegFormer variant's guard,
    eval_wsi_segmentation_gtcs.py:169-170)."""
    out = []
    for x_ind in range(slide_w // ws + 1):
        xmin = x_ind * ws
        xmax = slide_w if x_ind == slide_w // ws else (x_ind + 1) * ws
        if xmax > slide_w:
            continue
        for y_ind in range(slide_h // ws + 1):
            ymin = y_ind * ws
            ymax = slide_h if y_ind == slide_h // ws else (y_ind + 1) * ws
            if ymax > slide_w:          # sic: eval_wsi_segmentation.py:194
                continue
            out.append((xmin, ymin, xmax, ymax))
    return out


def check_overlap(gt: Sequence[float], ca: Sequence[float]) -> float:
    """annotation_handler.py:74-105 rectangle IoU (0.0 when the rectangles only touch)."""
    dx = min(ca[2], gt[2]) - max(ca[0], gt[0])
    dy = min(ca[3], gt[3]) - max(ca[1], gt[1])
    overlap = 0.0
    score = 0.0
    if dx > 0 and dy > 0:
        overlap = dx * dy
    if overlap > 0:
        area_ca = (ca[2] - ca[0]) * (ca[3] - ca[1])
        area_gt = (gt[2] - gt[0]) * (gt[3] - gt[1])
        score = overlap / (area_ca + area_gt - overlap)
    return score


# ------------------------------------------------------------------------------------------
# T3 -- paste boxes into a window with element-wise max
# ------------------------------------------------------------------------------------------
def overlay_window(boxes: Sequence[Sequence[float]], masks: Sequence[np.ndarray],
                   xmin: int, ymin: int, xmax: int, ymax: int) -> np.ndarray:
    """eval_wsi_segmentation.py:259-316 for the prediction case (times=1, margin 0):
    for every box whose rectangle IoU with the window is > 0, paste its class map on the union
    canvas (:301-307), slice the window back out and merge with np.max (:311-312)."""
    window = np.zeros((ymax - ymin, xmax - xmin), dtype=int)
    for box, m in zip(boxes, masks):
        b = [int(box[0]), int(box[1]), int(box[2]), int(box[3])]            # :262-266 (margin 0)
        if check_overlap([xmin, ymin, xmax, ymax], box[:4]) > 0.0:          # :268-269
            ux0, uy0 = min(xmin, b[0]), min(ymin, b[1])                      # :301-304
            ux1, uy1 = max(xmax, b[2]), max(ymax, b[3])
            canvas = np.zeros((int(uy1 - uy0), int(ux1 - ux0)), dtype=int)   # :305
            canvas[b[1] - uy0:b[3] - uy0, b[0] - ux0:b[2] - ux0] = m         # :307
            window = np.max(np.asarray((window, canvas[ymin - uy0:ymax - uy0, xmin - ux0:xmax - ux0]),
                                       dtype=int), axis=0)                   # :311-312
    return window


# ------------------------------------------------------------------------------------------
# T4 -- /8 nearest down-sample + paste
# ------------------------------------------------------------------------------------------
def nearest_index(dst_n: int, src_n: int) -> np.ndarray:
    """cv2.resize(..., INTER_NEAREST) source index (A11 / T4): sx = min(floor(dx * (src/dst)), src-1),
    the scale evaluated in double like OpenCV's resizeNN (verified against cv2 in tests)."""
    if dst_n == 0:
        return np.zeros((0,), np.int64)
    scale = float(src_n) / float(dst_n)        # OpenCV: ifx = 1. / (dst/src) computed as double
    idx = np.floor(np.arange(dst_n, dtype=np.float64) * scale).astype(np.int64)
    return np.minimum(idx, src_n - 1)


def resize_nearest(a: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    return a[nearest_index(dst_h, a.shape[0])][:, nearest_index(dst_w, a.shape[1])]


def stitch_slide(boxes, masks, slide_w: int, slide_h: int, ws: int):
    """Composition of T2 + T3 + T4 without the visualisation (palette / addWeighted,
    eval_wsi_segmentation.py:231-235 are out of scope): returns
      level0 [H,W] u8   -- every non-skipped window's `overlay` result at its place (zeros elsewhere)
      ds8    [int(H/8), int(W/8)] u8 -- generate_whole_img's label path: per window nearest resize to
             (int(w/8), int(h/8)) (:228) pasted at [ymin//8:ymax//8, xmin//8:xmax//8] (:236-240)."""
    level0 = np.zeros((slide_h, slide_w), np.uint8)
    ds8 = np.zeros((int(slide_h / MAGNIFICATION), int(slide_w / MAGNIFICATION)), np.uint8)   # :370
    for (xmin, ymin, xmax, ymax) in stitch_windows(slide_w, slide_h, ws):
        w, h = xmax - xmin, ymax - ymin
        if w == 0 or h == 0:
            continue
        win = overlay_window(boxes, masks, xmin, ymin, xmax, ymax)
        level0[ymin:ymax, xmin:xmax] = win
        small = resize_nearest(win, int(w / MAGNIFICATION), int(h / MAGNIFICATION))
        ds8[ymin // MAGNIFICATION:ymax // MAGNIFICATION, xmin // MAGNIFICATION:xmax // MAGNIFICATION] = small
    return level0, ds8


# ------------------------------------------------------------------------------------------
# IoU (module/common/IOUEval.py)
# ------------------------------------------------------------------------------------------
def fast_hist(gth: np.ndarray, pred: np.ndarray, n: int) -> np.ndarray:
    """IOUEval.py:19-21: rows = ground truth, columns = prediction."""
    a = gth.reshape(-1).astype(np.int64)
    b = pred.reshape(-1).astype(np.int64)
    k = (a >= 0) & (a < n)
    return np.bincount(n * a[k] + b[k], minlength=n * n).reshape(n, n)


def metric_right(hist: np.ndarray):
    """IOUEval.py:63-69 getMetricRight on an accumulated histogram."""
    eps = 0.00000001
    d = np.diag(hist)
    overall = d.sum() / (hist.sum() + eps)
    per_acc = d / (hist.sum(1) + eps)
    per_iou = d / (hist.sum(1) + hist.sum(0) - d + eps)
    return overall, per_acc, per_iou, np.nanmean(per_iou)
