"""Whole-slide tiling and tile -> slide stitching around the ESPNet forward.

Host side: the integer geometry of the reference's sliding-window tiler
(`module/faster-rcnn/detect_glomus_test.py:264-304`) and of the stitcher's window loop
(`module/espnet/test/eval_wsi_segmentation.py:180-195, 225-240`), evaluated with the same Python
float / int semantics so the indices are bit-exact.  Device side: tiles are read straight out of the
resident slide by the stem kernel, and the class maps are merged by the scatter / gather stitch kernels
of libespnet_b200.so.  Tile rows shard across ranks with no collective in the forward; the only
exchange is the final max-reduce of the slide masks to rank 0.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

MAGNIFICATION = 8   # eval_wsi_segmentation.py:22


@dataclass
class TileGrid:
    n_x: int
    n_y: int
    win_x: int
    win_y: int
    stride_x: int
    stride_y: int

    @property
    def count(self) -> int:
        return self.n_x * self.n_y

    def origins(self, row0: int = 0, rows: Optional[int] = None) -> np.ndarray:
        """(x_start, y_start) int32 [N,2] of the tiles of rows [row0,row0+rows), row-major
        `for j: for i:` like scan_region (detect_glomus_test.py:268-271)."""
        rows = self.n_y - row0 if rows is None else rows
        jj, ii = np.meshgrid(np.arange(row0, row0 + rows), np.arange(self.n_x), indexing="ij")
        return np.stack([ii.reshape(-1) * self.stride_x, jj.reshape(-1) * self.stride_y], 1).astype(np.int32)


def tile_grid(slide_w: int, slide_h: int, std_size: float = 512, mpp_x: float = 1.0, mpp_y: float = 1.0,
              overlap: float = 0.1, downsample: float = 1.0) -> TileGrid:
    """calc_window_size + the stride lines of scan_region (detect_glomus_test.py:286-304, 264-266)."""
    wx_org = float(std_size) / mpp_x
    wy_org = float(std_size) / mpp_y
    return TileGrid(
        n_x=int(math.ceil(slide_w / wx_org / (1.0 - overlap))),
        n_y=int(math.ceil(slide_h / wy_org / (1.0 - overlap))),
        win_x=int(math.ceil(wx_org / downsample)),
        win_y=int(math.ceil(wy_org / downsample)),
        stride_x=int(wx_org * (1.0 - overlap)),
        stride_y=int(wy_org * (1.0 - overlap)),
    )


def select_level(objective_power: float, level_downsamples: Sequence[float]) -> Tuple[int, float]:
    """detect_glomus_test.py:255-262: the first pyramid level at <= 5x magnification; when none qualifies the reference keeps
    its defaults target_level = 3, slide_downsample = 8.0."""
    target_level, downsample = 3, 8.0
    for level, ds in enumerate(level_downsamples):
        if objective_power / ds <= 5.0:
            target_level, downsample = level, level_downsamples[level]
            break
    return target_level, downsample


def stitch_y_limit(slide_w: int, slide_h: int, ws: int) -> int:
    """Rows >= this value are never written by the reference's window loop: windows whose ymax exceeds the
    slide WIDTH are skipped (`if ymax > slide_width: continue`, eval_wsi_segmentation.py:194, sic)."""
    lim = 0
    for y_ind in range(slide_h // ws + 1):
        ymax = slide_h if y_ind == slide_h // ws else (y_ind + 1) * ws
        if ymax > slide_w:
            continue
        lim = max(lim, ymax)
    return lim


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous band of tile rows for `rank`: (row0, rows)."""
    base, rem = divmod(n_rows, world)
    row0 = rank * base + min(rank, rem)
    return row0, base + (1 if rank < rem else 0)


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def stitch_boxes(slide_mask: torch.Tensor, boxes: Sequence[Sequence[float]], masks: Sequence[torch.Tensor], ws: int = 2400):
    """T3 for arbitrary boxes (eval_wsi_segmentation.py:259-316): every class map is max-merged into the
    level-0 slide mask at its box (int() truncated like :262-266).  `masks[i]` is uint8 [y1-y0, x1-x0] on the
    slide's device.  slide_mask: zero-initialised uint8 [SH,SW]."""
    sh, sw = slide_mask.shape
    dev = slide_mask.device
    ib = np.array([[int(b[0]), int(b[1]), int(b[2]), int(b[3])] for b in boxes], np.int32).reshape(-1, 4)
    sizes = [(int(b[3] - b[1])) * (int(b[2] - b[0])) for b in ib]
    for m, b, s in zip(masks, ib, sizes):
        if m.dtype != torch.uint8 or m.numel() != s:
            raise RuntimeError("class map of box %s has the wrong size/dtype" % (b.tolist(),))
    offs = np.zeros(len(ib), np.int64)
    if len(ib):
        offs[1:] = np.cumsum(sizes)[:-1]
    flat = torch.cat([m.reshape(-1) for m in masks]) if len(masks) else torch.zeros(0, dtype=torch.uint8, device=dev)
    d_boxes = torch.from_numpy(ib).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    rc = _lib.lib().espnet_stitch_boxes(slide_mask.data_ptr(), sh, sw, stitch_y_limit(sw, sh, ws), d_boxes.data_ptr(),
                                        d_offs.data_ptr(), flat.data_ptr(), len(ib), _stream(dev))
    _lib.check(rc, None, "espnet_stitch_boxes")
    return slide_mask


def stitch_grid(slide_mask: torch.Tensor, tile_masks: torch.Tensor, grid: TileGrid, row0: int, rows: int, ws: int = 2400):
    """T3 for the regular tile grid, gather form.  tile_masks: uint8 [rows*n_x, win_y, win_x]."""
    sh, sw = slide_mask.shape
    rc = _lib.lib().espnet_stitch_grid(slide_mask.data_ptr(), sh, sw, stitch_y_limit(sw, sh, ws), tile_masks.data_ptr(),
                                       grid.n_x, grid.n_y, grid.win_x, grid.win_y, grid.stride_x, grid.stride_y, row0, rows,
                                       _stream(slide_mask.device))
    _lib.check(rc, None, "espnet_stitch_grid")
    return slide_mask


def ds8_luts(slide_w: int, slide_h: int, ws: int) -> Tuple[np.ndarray, np.ndarray]:
    """Source-row / source-column LUTs of generate_whole_img's label path (eval_wsi_segmentation.py:225-240),
    computed in double inside the library (espnet_ds8_lut)."""
    if ws % MAGNIFICATION:
        raise RuntimeError("window_size must be a multiple of 8: the reference's paste at [xmin//8:xmax//8] only "
                           "matches its int(w/8) resize then (eval_wsi_segmentation.py:228,236-240)")
    dw, dh = int(slide_w / MAGNIFICATION), int(slide_h / MAGNIFICATION)
    xs = np.empty(dw, np.int32)
    ys = np.empty(dh, np.int32)
    _lib.check(_lib.lib().espnet_ds8_lut(slide_w, ws, slide_w, xs.ctypes.data, dw), None, "espnet_ds8_lut(x)")
    _lib.check(_lib.lib().espnet_ds8_lut(slide_h, ws, slide_w, ys.ctypes.data, dh), None, "espnet_ds8_lut(y)")
    return ys, xs


def downsample8(level0: torch.Tensor, ws: int = 2400) -> torch.Tensor:
    """T4: the /8 label image the reference pastes window by window (before palette / blending)."""
    sh, sw = level0.shape
    ys, xs = ds8_luts(sw, sh, ws)
    dev = level0.device
    d_ys, d_xs = torch.from_numpy(ys).to(dev), torch.from_numpy(xs).to(dev)
    out = torch.empty((len(ys), len(xs)), dtype=torch.uint8, device=dev)
    rc = _lib.lib().espnet_downsample_lut(level0.data_ptr(), sh, sw, out.data_ptr(), len(ys), len(xs), d_ys.data_ptr(),
                                          d_xs.data_ptr(), _stream(dev))
    _lib.check(rc, None, "espnet_downsample_lut")
    return out


def segment_slide(model, slide_u8: torch.Tensor, mean, std, std_size: float = 512, mpp: float = 1.0, overlap: float = 0.1,
                  downsample: float = 1.0, ws: int = 2400, batch: int = 256, rank: int = 0, world: int = 1,
                  reduce_to_rank0: bool = True):
    """Overlapping-tile WSI segmentation (BASELINE.json config 4): T1 tiles -> ESPNet forward + arg-max per
    tile -> T3 max-merge -> T4 /8 mask.  `slide_u8` is the resident BGR slide uint8 [SH,SW,3] (every rank holds
    it, or at least its band).  With world > 1 the tile rows are sharded across ranks; the forward has no
    collective, the only exchange is one max-reduce of the level-0 mask to rank 0.
    Returns (level0 uint8 [SH,SW], ds8 uint8 [int(SH/8), int(SW/8)], n_local_tiles)."""
    if downsample != 1.0:
        raise RuntimeError("only level-0 tiling (downsample 1) is wired to the resident-slide reader")
    sh, sw = int(slide_u8.shape[0]), int(slide_u8.shape[1])
    dev = slide_u8.device
    grid = tile_grid(sw, sh, std_size, mpp, mpp, overlap, downsample)
    if grid.win_x % 8 or grid.win_y % 8:
        raise RuntimeError("tile size %dx%d is not a multiple of 8 (ESPNet needs it, Model.py cat at :373)" % (grid.win_x, grid.win_y))
    row0, rows = shard_rows(grid.n_y, rank, world)
    level0 = torch.zeros((sh, sw), dtype=torch.uint8, device=dev)
    n_local = rows * grid.n_x
    if n_local:
        origins = torch.from_numpy(grid.origins(row0, rows)).to(dev)
        masks = torch.empty((n_local, grid.win_y, grid.win_x), dtype=torch.uint8, device=dev)
        for s in range(0, n_local, batch):
            e = min(s + batch, n_local)
            model.segment_tiles(slide_u8, origins[s:e], grid.win_y, grid.win_x, mean, std, out=masks[s:e])
        stitch_grid(level0, masks, grid, row0, rows, ws)
    if world > 1 and reduce_to_rank0:
        import torch.distributed as dist
        dist.reduce(level0, dst=0, op=dist.ReduceOp.MAX)
    ds8 = downsample8(level0, ws) if (rank == 0 or not reduce_to_rank0) else None
    return level0, ds8, n_local
