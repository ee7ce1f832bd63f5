"""Seeded case definitions shared by `tests/golden/make_wsi_golden.py` (which feeds them to the REAL reference
scripts) and by the tests (which feed the same inputs to the oracle and to the CUDA path).  Only inputs live
here; every expected output comes from the committed fixtures."""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------------
# T1: GlomusDetector.scan_region (detect_glomus_test.py:236-304)
#   (slide_w, slide_h, STD_SIZE um, mpp_x, mpp_y, overlap, objective_power, level_downsamples)
# ---------------------------------------------------------------------------------------------
T1_CASES = [
    (40000, 30000, 512, 1.0, 1.0, 0.1, 5, (1.0, 2.0, 4.0)),                       # BASELINE configs[3]
    (40000, 30000, 512, 1.0, 1.0, 0.5, 5, (1.0, 2.0, 4.0)),
    (4096, 3072, 512, 1.0, 1.0, 0.1, 5, (1.0, 4.0)),
    (98304, 71680, 2000, 0.2277, 0.2277, 0.1, 40, (1.0, 2.0, 4.0, 8.0, 16.0, 32.0)),   # example/README.md:34-35
    (51234, 40321, 500, 0.2277, 0.2301, 0.35, 40, (1.0, 2.0, 4.0, 8.0, 16.0)),
    (300, 200, 512, 1.0, 1.0, 0.1, 5, (1.0,)),                                    # slide smaller than one window
    (53248, 23040, 2000, 0.2277, 0.2277, 0.1, 20, (1.0, 2.0, 4.000123, 8.0002)),  # non-integer level downsample
    (20000, 10000, 700, 0.4554, 0.4554, 0.25, 40, (1.0, 2.0, 4.0)),               # no level reaches 5x: level 3 / ds 8.0 default
    (12345, 6789, 333, 0.5, 0.25, 0.0, 10, (1.0, 2.0, 4.0)),                      # zero overlap, anisotropic mpp
    (420, 300, 96, 1.0, 1.0, 0.25, 5, (1.0,)),                                    # the size the GPU slide-reader test uses
]

# ---------------------------------------------------------------------------------------------
# T2/T3/T4 (+ f3 render): Generate_Segmentation_Gt.generate_pred_wsi (eval_wsi_segmentation.py:359-394)
#   name -> (slide_w, slide_h, window_size, n_boxes, box side range, seed, cityscapes ids?, expect crash?)
# ---------------------------------------------------------------------------------------------
STITCH_CASES = {
    "wide":      dict(sw=2000, sh=1500, ws=640, n=28, side=(60, 520), seed=11, city=False),
    "tall":      dict(sw=900, sh=2100, ws=400, n=30, side=(40, 380), seed=12, city=False),     # ymax > slide_width quirk
    "city":      dict(sw=1333, sh=1111, ws=480, n=22, side=(30, 400), seed=13, city=True),     # ids 7/8/11/12/13 relabelled
    "default":   dict(sw=5000, sh=3700, ws=2400, n=40, side=(200, 1100), seed=14, city=False),  # the reference's default window
    "dense":     dict(sw=700, sh=650, ws=160, n=60, side=(8, 200), seed=15, city=False),       # heavy box overlap
    "divisible": dict(sw=1280, sh=700, ws=640, n=6, side=(50, 300), seed=16, city=False),      # W % ws == 0: reference crashes
}

CITY_IDS = np.array([7, 8, 11, 12, 13], np.uint8)     # VisualizeResults_iou.py:54-81 <-> eval_wsi_segmentation.py:49-55


def blobby_mask(rng, h, w, noisy):
    """Class map 0..4 [h,w]: 8-px blobs (compressible fixtures), per-pixel noise for a few boxes."""
    if noisy:
        return rng.integers(0, 5, (h, w)).astype(np.uint8)
    coarse = rng.integers(0, 5, (h // 8 + 2, w // 8 + 2)).astype(np.uint8)
    oy, ox = int(rng.integers(0, 8)), int(rng.integers(0, 8))
    return np.kron(coarse, np.ones((8, 8), np.uint8))[oy:oy + h, ox:ox + w].copy()


def stitch_inputs(name):
    """(boxes [[xmin,ymin,xmax,ymax,conf]] with integer coordinates like the merged-detection CSV
    (eval_wsi_segmentation.py:330), class maps uint8 [ymax-ymin, xmax-xmin] with values 0..4)."""
    c = STITCH_CASES[name]
    rng = np.random.default_rng(c["seed"])
    boxes, masks, names = [], [], set()
    while len(boxes) < c["n"]:
        w, h = int(rng.integers(*c["side"])), int(rng.integers(*c["side"]))
        x0 = int(rng.integers(-c["side"][0], c["sw"] - 4))
        y0 = int(rng.integers(-c["side"][0], c["sh"] - 4))
        if len(boxes) == 0:                      # one box exactly touching a window corner from outside: IoU 0 with it
            x0, y0 = c["ws"], c["ws"]
        if len(boxes) == 1:                      # one box ending exactly on a window edge
            x0, y0 = c["ws"] - w, max(0, c["ws"] // 2 - h)
        key = (int(x0 / 8), int(y0 / 8), int((x0 + w) / 8), int((y0 + h) / 8))    # the JSON search name (:272)
        if key in names:
            continue
        names.add(key)
        boxes.append([x0, y0, x0 + w, y0 + h, float(np.round(rng.uniform(0.5, 1.0), 3))])
        masks.append(blobby_mask(rng, h, w, noisy=(len(boxes) % 7 == 0)))
    return boxes, masks


def slide_rgb(name):
    """Synthetic level-0 slide, uint8 RGB [sh, sw, 3] (what openslide's read_region hands back, alpha dropped)."""
    c = STITCH_CASES[name]
    rng = np.random.default_rng(c["seed"] + 1000)
    return rng.integers(0, 256, (c["sh"], c["sw"], 3), dtype=np.uint8)


# ---------------------------------------------------------------------------------------------
# P0 / f2 / A10 / A11 / f4: evaluateModel (VisualizeResults_iou.py:84-156) on crops whose size differs from
# the network size.  (crop_h, crop_w, inWidth, inHeight, fold, distribution, seed)
# ---------------------------------------------------------------------------------------------
FRONTEND_CASES = [
    (300, 420, 256, 128, 1, "D2", 21),      # down-scale, W x H = 2:1 like the reference default 1024x512
    (97, 131, 160, 96, 3, "D2", 22),        # up-scale, odd crop size
    (128, 256, 256, 128, 5, "D1", 23),      # identity size (cv2.resize returns a copy)
    (511, 389, 128, 192, 2, "D2", 24),      # portrait network size
]

# rectangles for AnnotationHandler.check_overlap (annotation_handler.py:74-105)
def overlap_pairs():
    rng = np.random.default_rng(5)
    pairs = [([0, 0, 10, 10], [10, 0, 20, 10]), ([0, 0, 10, 10], [0, 10, 10, 20]), ([0, 0, 10, 10], [2, 2, 8, 8]),
             ([0, 0, 10, 10], [0, 0, 10, 10]), ([0, 0, 10, 10], [9, 9, 30, 30]), ([5, 5, 6, 6], [0, 0, 100, 100]),
             ([0.5, 0.25, 10.75, 9.125], [10.5, 0.0, 12.0, 3.0]), ([-5, -5, 5, 5], [0, 0, 3, 3])]
    for _ in range(200):
        a = rng.integers(-50, 400, 2); b = rng.integers(-50, 400, 2)
        ra = [int(a[0]), int(a[1]), int(a[0] + rng.integers(1, 200)), int(a[1] + rng.integers(1, 200))]
        rb = [float(b[0]) + 0.5 * int(rng.integers(0, 2)), float(b[1]), float(b[0] + rng.integers(1, 200)), float(b[1] + rng.integers(1, 200))]
        pairs.append((ra, rb))
    return pairs
