"""Crop front-end and render around the forward, on the GPU through the C ABI:
  preprocess_resize      VisualizeResults_iou.py:107-119  (normalise, cv2.resize INTER_LINEAR, /255, HWC->CHW)
  resize_mask_nearest    VisualizeResults_iou.py:129      (class map back to the crop size, INTER_NEAREST)
  colorize / overlay     VisualizeResults_iou.py:139-147  (palette colour map, cv2.addWeighted(img, .4, map, .6, 0))
  render_slide_ds8       eval_wsi_segmentation.py:225-240 (the /8 rendered slide, all windows at once)
  class_pixel_counts     VisualizeResults_iou.py:151-155
Index tables are computed on the host by the library (double / float arithmetic identical to OpenCV's), so indices are
bit-exact; u8 results are bit-exact, the bilinear float result equals OpenCV's generic (non-IPP) code path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .wsi import ds8_luts

# eval_wsi_segmentation.py:23-47 / VisualizeResults_iou.py:22-46 (rows are r, g, b)
PALLETE = [[0, 0, 0], [255, 0, 0], [0, 184, 0], [255, 255, 0], [0, 0, 255], [128, 64, 128], [244, 35, 232], [70, 70, 70],
           [102, 102, 156], [190, 153, 153], [153, 153, 153], [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152],
           [70, 130, 180], [220, 20, 60], [255, 0, 0], [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230],
           [119, 11, 32], [0, 0, 0]]


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def bilinear_lut(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    idx, wgt = np.empty(dst, np.int32), np.empty(dst, np.float32)
    _lib.check(_lib.lib().espnet_bilinear_lut(src, dst, idx.ctypes.data, wgt.ctypes.data), None, "espnet_bilinear_lut")
    return idx, wgt


def nearest_lut(src: int, dst: int) -> np.ndarray:
    idx = np.empty(dst, np.int32)
    _lib.check(_lib.lib().espnet_nearest_lut(src, dst, idx.ctypes.data), None, "espnet_nearest_lut")
    return idx


def _palette(dev) -> torch.Tensor:
    return torch.tensor(PALLETE, dtype=torch.uint8, device=dev).contiguous()


def preprocess_resize(crops_u8: torch.Tensor, mean: Sequence[float], std: Sequence[float], width: int, height: int) -> torch.Tensor:
    """crops_u8: CUDA uint8 [B,h,w,3] BGR -> fp32 [B,3,height,width], what `model(img_variable)` receives (:107-123)."""
    if crops_u8.dtype != torch.uint8 or crops_u8.dim() != 4 or crops_u8.shape[-1] != 3 or not crops_u8.is_cuda:
        raise RuntimeError("preprocess_resize() wants a CUDA uint8 [B,h,w,3] BGR tensor")
    crops_u8 = crops_u8.contiguous()
    B, h, w, _ = crops_u8.shape
    dev = crops_u8.device
    xs, xf = bilinear_lut(w, width)
    ys, yf = bilinear_lut(h, height)
    d = [torch.from_numpy(a).to(dev) for a in (xs, xf, ys, yf)]
    out = torch.empty((B, 3, height, width), dtype=torch.float32, device=dev)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    rc = _lib.lib().espnet_preprocess_resize(crops_u8.data_ptr(), B, h, w, m, s, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(),
                                             d[3].data_ptr(), out.data_ptr(), height, width, _stream(dev))
    _lib.check(rc, None, "espnet_preprocess_resize")
    return out


def crop_regions(boxes: Sequence[Sequence[float]]):
    """make_seg_data.py:347-361: per detected box the level-0 read_region arguments (x, y, w, h) and the crop's file
    stem `xmin{}_ymin{}_xmax{}_ymax{}` in /8 coordinates (:359) -- the name overlay() later searches for
    (eval_wsi_segmentation.py:272)."""
    regions = [(b[0], b[1], b[2] - b[0], b[3] - b[1]) for b in boxes]
    names = ["xmin{}_ymin{}_xmax{}_ymax{}".format(int(b[0] / 8), int(b[1] / 8), int(b[2] / 8), int(b[3] / 8)) for b in boxes]
    return regions, names


def preprocess_boxes(slide_u8: torch.Tensor, boxes: Sequence[Sequence[float]], mean: Sequence[float], std: Sequence[float],
                     width: int = 1024, height: int = 512) -> torch.Tensor:
    """Box -> crop -> network input in one kernel: what make_seg_data.output_org_files (:347-361) cuts out of the slide and
    VisualizeResults_iou.py:103-119 then normalises and resizes to inWidth x inHeight (defaults 1024 x 512, :297-298).
    slide_u8: resident CUDA uint8 [SH,SW,3] in BGR (cv2.imread order); boxes: [[xmin,ymin,xmax,ymax,...]] level-0 px, may
    overhang the slide (zero padding).  Returns fp32 [B,3,height,width]."""
    if slide_u8.dtype != torch.uint8 or slide_u8.dim() != 3 or slide_u8.shape[-1] != 3 or not slide_u8.is_cuda or not slide_u8.is_contiguous():
        raise RuntimeError("preprocess_boxes() wants a contiguous CUDA uint8 [SH,SW,3] BGR slide")
    ib = np.array([[int(b[0]), int(b[1]), int(b[2]), int(b[3])] for b in boxes], np.int32).reshape(-1, 4)
    if len(ib) == 0 or (ib[:, 2] <= ib[:, 0]).any() or (ib[:, 3] <= ib[:, 1]).any():
        raise RuntimeError("preprocess_boxes() needs at least one box and positive box sizes")
    B = len(ib)
    dev = slide_u8.device
    xs, xf = np.empty((B, width), np.int32), np.empty((B, width), np.float32)
    ys, yf = np.empty((B, height), np.int32), np.empty((B, height), np.float32)
    cache = {}
    for i, (x0, y0, x1, y1) in enumerate(ib):
        for n, dst, (ia, fa) in ((int(x1 - x0), width, (xs, xf)), (int(y1 - y0), height, (ys, yf))):
            if (n, dst) not in cache:
                cache[(n, dst)] = bilinear_lut(n, dst)
            ia[i], fa[i] = cache[(n, dst)]
    d = [torch.from_numpy(a).to(dev) for a in (ib, xs, xf, ys, yf)]
    out = torch.empty((B, 3, height, width), dtype=torch.float32, device=dev)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(dev):
        rc = _lib.lib().espnet_preprocess_resize_boxes(slide_u8.data_ptr(), int(slide_u8.shape[0]), int(slide_u8.shape[1]), d[0].data_ptr(), B, m, s,
                                                       d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(), out.data_ptr(),
                                                       height, width, _stream(dev))
    _lib.check(rc, None, "espnet_preprocess_resize_boxes")
    return out


def resize_mask_nearest(masks: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """CUDA uint8 [B,H,W] -> [B,height,width] with cv2 INTER_NEAREST indices (:129)."""
    if masks.dtype != torch.uint8 or masks.dim() != 3 or not masks.is_cuda:
        raise RuntimeError("resize_mask_nearest() wants a CUDA uint8 [B,H,W] tensor")
    masks = masks.contiguous()
    B, sh, sw = masks.shape
    dev = masks.device
    ys = torch.from_numpy(nearest_lut(sh, height)).to(dev)
    xs = torch.from_numpy(nearest_lut(sw, width)).to(dev)
    out = torch.empty((B, height, width), dtype=torch.uint8, device=dev)
    rc = _lib.lib().espnet_resize_nearest_u8(masks.data_ptr(), B, sh, sw, out.data_ptr(), height, width, ys.data_ptr(), xs.data_ptr(), _stream(dev))
    _lib.check(rc, None, "espnet_resize_nearest_u8")
    return out


def colorize_overlay(label: torch.Tensor, img_u8: Optional[torch.Tensor] = None):
    """label: CUDA uint8 [...]; img_u8: CUDA uint8 [...,3] of the same pixels (BGR like cv2.imread).  Returns
    (colour map [...,3], overlay [...,3] or None) -- classMap_numpy_color and cv2.addWeighted(img, 0.4, map, 0.6, 0) (:139-147)."""
    label = label.contiguous()
    dev = label.device
    color = torch.empty(tuple(label.shape) + (3,), dtype=torch.uint8, device=dev)
    over = None
    if img_u8 is not None:
        img_u8 = img_u8.contiguous()
        if tuple(img_u8.shape) != tuple(color.shape):
            raise RuntimeError("image and label shapes differ")
        over = torch.empty_like(color)
    pal = _palette(dev)
    rc = _lib.lib().espnet_palette_overlay(img_u8.data_ptr() if img_u8 is not None else None, label.data_ptr(), label.numel(), pal.data_ptr(),
                                           len(PALLETE), color.data_ptr(), over.data_ptr() if over is not None else None, _stream(dev))
    _lib.check(rc, None, "espnet_palette_overlay")
    return color, over


def render_slide_ds8(slide_u8: torch.Tensor, level0_mask: torch.Tensor, ws: int = 2400) -> torch.Tensor:
    """The whole `*_pred.jpg` canvas of generate_pred_wsi (eval_wsi_segmentation.py:359-394): every window's /8 nearest slide
    pixels blended with its /8 nearest palette map; windows the reference's loop skips stay black."""
    sh, sw = level0_mask.shape
    dev = level0_mask.device
    ys, xs = ds8_luts(sw, sh, ws)
    d_ys, d_xs = torch.from_numpy(ys).to(dev), torch.from_numpy(xs).to(dev)
    out = torch.empty((len(ys), len(xs), 3), dtype=torch.uint8, device=dev)
    pal = _palette(dev)
    rc = _lib.lib().espnet_render_ds8(slide_u8.contiguous().data_ptr(), level0_mask.contiguous().data_ptr(), sh, sw, pal.data_ptr(), len(PALLETE),
                                      out.data_ptr(), len(ys), len(xs), d_ys.data_ptr(), d_xs.data_ptr(), _stream(dev))
    _lib.check(rc, None, "espnet_render_ds8")
    return out


def class_pixel_counts(masks: torch.Tensor, n_classes: int = 5) -> torch.Tensor:
    """int64 [B, n_classes]: background / glomeruli / crescent / sclerosis / mesangium pixel counts (:151-155)."""
    masks = masks.contiguous()
    B = masks.shape[0]
    counts = torch.zeros((B, n_classes), dtype=torch.int64, device=masks.device)
    rc = _lib.lib().espnet_class_counts(masks.data_ptr(), B, masks[0].numel(), n_classes, counts.data_ptr(), _stream(masks.device))
    _lib.check(rc, None, "espnet_class_counts")
    return counts
