"""Whole-slide tiling and tile -> slide stitching around the ESPNet forward.

Host side: the integer geometry of the reference's sliding-window tiler
(`module/faster-rcnn/detect_glomus_test.py:264-304`) and of the stitcher's window loop
(`module/espnet/test/eval_wsi_segmentation.py:180-195, 225-240`), evaluated with the same Python
float / int semantics so the indices are bit-exact.  Device side: tiles are read straight out of the
resident slide by the stem kernel, and the class maps are merged by the scatter / gather stitch kernels
of libespnet_b200.so.  Tile rows shard across ranks with no collective in the forward; every rank holds and
stitches only its own BAND of slide rows, and the only exchange is the final gather of the band masks on rank 0
(point-to-point, placed straight into the slide mask; only the rows two bands share -- the tile overlap -- go
through an element-wise max).
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

    def origins_range(self, k0: int, k1: int) -> np.ndarray:
        """(x_start, y_start) int32 of the tiles with row-major index k0 <= k < k1 (k = j * n_x + i, scan_region's order)."""
        k = np.arange(k0, k1)
        return np.stack([(k % self.n_x) * self.stride_x, (k // self.n_x) * self.stride_y], 1).astype(np.int32)


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


def band_rows(grid: TileGrid, slide_h: int, rank: int, world: int) -> Tuple[int, int, int, int]:
    """(tile_row0, tile_rows, y0, y1): whole tile rows for `rank` and the slide rows [y0, y1) their tiles cover
    (clipped to the slide).  Adjacent bands share win_y - stride_y rows (the tile overlap)."""
    row0, rows = shard_rows(grid.n_y, rank, world)
    if rows == 0:
        return row0, 0, 0, 0
    y0 = row0 * grid.stride_y
    y1 = min(slide_h, (row0 + rows - 1) * grid.stride_y + grid.win_y)
    return row0, rows, min(y0, slide_h), max(y1, min(y0, slide_h))


def band_tiles(grid: TileGrid, slide_h: int, rank: int, world: int, balance: str = "tiles") -> Tuple[int, int, int, int]:
    """(k0, k1, y0, y1): the row-major tile index range of `rank` and the slide rows [y0, y1) those tiles touch.
    balance="tiles" (default) splits the TILES evenly -- a range may start / end inside a tile row, the two neighbours then both
    touch that row's slide rows and the stitch merges them; balance="rows" hands out whole tile rows (band_rows)."""
    if balance == "rows":
        row0, rows, y0, y1 = band_rows(grid, slide_h, rank, world)
        return row0 * grid.n_x, (row0 + rows) * grid.n_x, y0, y1
    if balance != "tiles":
        raise ValueError(balance)
    k0, n = shard_rows(grid.count, rank, world)        # the same contiguous split, over tiles instead of rows
    k1 = k0 + n
    if n == 0:
        return k0, k1, 0, 0
    y0 = min((k0 // grid.n_x) * grid.stride_y, slide_h)
    y1 = min(slide_h, ((k1 - 1) // grid.n_x) * grid.stride_y + grid.win_y)
    return k0, k1, y0, max(y1, y0)


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _check_u8(t: torch.Tensor, what: str, ndim: int, dev=None):
    if not isinstance(t, torch.Tensor) or t.dtype != torch.uint8 or t.dim() != ndim or not t.is_cuda or not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous CUDA uint8 tensor with %d dimensions" % (what, ndim))
    if dev is not None and t.device != dev:
        raise RuntimeError("%s lives on %s, expected %s" % (what, t.device, dev))


def stitch_boxes(slide_mask: torch.Tensor, boxes: Sequence[Sequence[float]], masks: Sequence[torch.Tensor], ws: int = 2400):
    """T3 for arbitrary boxes (eval_wsi_segmentation.py:259-316): every class map is max-merged into the
    level-0 slide mask at its box (int() truncated like :262-266).  `masks[i]` is uint8 [y1-y0, x1-x0] on the
    slide's device.  slide_mask: zero-initialised uint8 [SH,SW]."""
    _check_u8(slide_mask, "slide_mask", 2)
    sh, sw = slide_mask.shape
    dev = slide_mask.device
    if slide_mask.data_ptr() % 4:
        raise RuntimeError("slide_mask must be 4-byte aligned (32-bit merge words)")
    ib = np.array([[int(b[0]), int(b[1]), int(b[2]), int(b[3])] for b in boxes], np.int32).reshape(-1, 4)
    sizes = [(int(b[3] - b[1])) * (int(b[2] - b[0])) for b in ib]
    for m, b, s in zip(masks, ib, sizes):
        if m.dtype != torch.uint8 or m.numel() != s or m.device != dev:
            raise RuntimeError("class map of box %s has the wrong size / dtype / device" % (b.tolist(),))
    offs = np.zeros(len(ib), np.int64)
    if len(ib):
        offs[1:] = np.cumsum(sizes)[:-1]
    flat = torch.cat([m.reshape(-1) for m in masks]) if len(masks) else torch.zeros(0, dtype=torch.uint8, device=dev)
    d_boxes = torch.from_numpy(ib).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().espnet_stitch_boxes(slide_mask.data_ptr(), sh, sw, stitch_y_limit(sw, sh, ws), d_boxes.data_ptr(),
                                            d_offs.data_ptr(), flat.data_ptr(), len(ib), _stream(dev))
    _lib.check(rc, None, "espnet_stitch_boxes")
    return slide_mask


def stitch_grid(slide_mask: torch.Tensor, tile_masks: torch.Tensor, grid: TileGrid, row0: int, rows: int, ws: int = 2400,
                band_y0: int = 0, slide_h: Optional[int] = None, overwrite: bool = False, tiles: Optional[Tuple[int, int]] = None):
    """T3 for the regular tile grid, gather form.  tile_masks: uint8 [n, win_y, win_x] = the tiles of rows [row0, row0 + rows)
    or, with tiles=(k0, k1), the row-major tile index range [k0, k1).  `slide_mask` is the whole level-0 mask [SH,SW] (default)
    or, with `slide_h` given, a band buffer holding slide rows [band_y0, band_y0 + its height).
    overwrite=True writes every pixel of the covered rows without reading the destination."""
    k0, k1 = tiles if tiles is not None else (row0 * grid.n_x, (row0 + rows) * grid.n_x)
    _check_u8(slide_mask, "slide_mask", 2)
    _check_u8(tile_masks, "tile_masks", 3, None if overwrite else slide_mask.device)
    if tuple(tile_masks.shape) != (k1 - k0, grid.win_y, grid.win_x):
        raise RuntimeError("tile_masks must be [n_tiles, win_y, win_x] = %s, got %s" % ((k1 - k0, grid.win_y, grid.win_x), tuple(tile_masks.shape)))
    bh, sw = slide_mask.shape
    sh = bh if slide_h is None else int(slide_h)
    dev = tile_masks.device                    # the kernel runs where the tiles are
    with torch.cuda.device(dev):
        rc = _lib.lib().espnet_stitch_grid_band(slide_mask.data_ptr(), band_y0, bh, sh, sw, stitch_y_limit(sw, sh, ws), tile_masks.data_ptr(),
                                                grid.n_x, grid.n_y, grid.win_x, grid.win_y, grid.stride_x, grid.stride_y, k0, k1,
                                                int(overwrite), _stream(dev))
    _lib.check(rc, None, "espnet_stitch_grid_band")
    return slide_mask


def max_merge_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst = max(dst, src) element-wise on CUDA uint8 tensors of equal size (the overlap-strip merge of the band gather)."""
    if dst.numel() != src.numel() or not dst.is_contiguous() or not src.is_contiguous() or dst.dtype != torch.uint8 or src.dtype != torch.uint8:
        raise RuntimeError("max_merge_ wants two contiguous uint8 tensors of equal size")
    with torch.cuda.device(dst.device):
        rc = _lib.lib().espnet_max_merge_u8(dst.data_ptr(), src.data_ptr(), dst.numel(), _stream(dst.device))
    _lib.check(rc, None, "espnet_max_merge_u8")
    return dst


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
    _check_u8(level0, "level0", 2)
    sh, sw = level0.shape
    ys, xs = ds8_luts(sw, sh, ws)
    dev = level0.device
    d_ys, d_xs = torch.from_numpy(ys).to(dev), torch.from_numpy(xs).to(dev)
    out = torch.empty((len(ys), len(xs)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().espnet_downsample_lut(level0.data_ptr(), sh, sw, out.data_ptr(), len(ys), len(xs), d_ys.data_ptr(),
                                              d_xs.data_ptr(), _stream(dev))
    _lib.check(rc, None, "espnet_downsample_lut")
    return out


def gather_bands(band: Optional[torch.Tensor], grid: TileGrid, slide_h: int, slide_w: int, rank: int, world: int,
                 merge=None, group=None, balance: str = "tiles") -> Tuple[Optional[torch.Tensor], dict]:
    """The one exchange of the multi-GPU WSI path: every rank hands its band mask (slide rows
    [y0_r, y1_r), `band_rows`) to rank 0.  Rank 0 allocates the level-0 mask, keeps its own band, receives the rows no earlier
    band covers STRAIGHT INTO their place in the slide mask, and only the rows a band shares with its predecessors (the tile
    overlap, win - stride rows) into a staging buffer that is max-merged (eval_wsi_segmentation.py:311-312).  Point-to-point
    (NCCL send/recv over NVLink, grouped), no reduction over the whole mask.  Returns (level0 on rank 0 else None, stats)."""
    import torch.distributed as dist
    merge = merge or (lambda d, s: max_merge_(d, s))
    bands = [band_tiles(grid, slide_h, r, world, balance) for r in range(world)]
    bands = [(k0, k1 - k0, y0, y1) for k0, k1, y0, y1 in bands]          # (first tile, tile count, y0, y1)
    stats = {"bytes_received": 0, "bytes_merged": 0}
    if rank != 0:
        _, rows, y0, y1 = bands[rank]
        if rows and y1 > y0:
            cov = max([b[3] for b in bands[:rank] if b[1]] + [0])
            split = min(max(cov, y0), y1) - y0                      # band rows [0, split) overlap earlier bands
            ops = []
            if split > 0:
                ops.append(dist.P2POp(dist.isend, band[:split], 0, group))
            if split < y1 - y0:
                ops.append(dist.P2POp(dist.isend, band[split:], 0, group))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return None, stats
    level0 = torch.zeros((slide_h, slide_w), dtype=torch.uint8, device=band.device if band is not None else None)
    _, rows, y0, y1 = bands[0]
    cov = 0
    if rows and y1 > y0:
        level0[y0:y1].copy_(band)
        cov = y1
    ops, strips = [], []
    for r in range(1, world):
        _, rows, y0, y1 = bands[r]
        if not rows or y1 <= y0:
            continue
        split = min(max(cov, y0), y1)
        if split > y0:
            st = torch.empty((split - y0, slide_w), dtype=torch.uint8, device=level0.device)
            strips.append((y0, split, st))
            ops.append(dist.P2POp(dist.irecv, st, r, group))
        if split < y1:
            ops.append(dist.P2POp(dist.irecv, level0[split:y1], r, group))      # zero-copy placement
        stats["bytes_received"] += (y1 - y0) * slide_w
        cov = max(cov, y1)
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for ya, yb, st in strips:
        merge(level0[ya:yb], st)
        stats["bytes_merged"] += (yb - ya) * slide_w
    return level0, stats


def band_plan(grid: TileGrid, slide_h: int, world: int, balance: str = "tiles"):
    """Per rank (y0, split, y1): band rows [y0, y1); rows [y0, split) are also touched by an earlier rank's band (they need the
    max-merge), rows [split, y1) are this rank's to place."""
    plan, cov = [], 0
    for r in range(world):
        k0, k1, y0, y1 = band_tiles(grid, slide_h, r, world, balance)
        if k1 <= k0 or y1 <= y0:
            plan.append((0, 0, 0))
            continue
        split = min(max(cov, y0), y1)
        plan.append((y0, split, y1))
        cov = max(cov, y1)
    return plan


class _DeviceBytes:
    """__cuda_array_interface__ view of raw device memory, so that torch can alias it without owning it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _stitch_raw(out_ptr: int, out_y0: int, out_rows: int, sh: int, sw: int, tile_masks: torch.Tensor, grid: TileGrid, k0: int, k1: int, ws: int):
    """overwrite-form band stitch of tiles [k0, k1) into a raw (possibly peer-mapped) pointer; runs on the tiles' device and stream"""
    dev = tile_masks.device
    with torch.cuda.device(dev):
        rc = _lib.lib().espnet_stitch_grid_band(out_ptr, out_y0, out_rows, sh, sw, stitch_y_limit(sw, sh, ws), tile_masks.data_ptr(), grid.n_x,
                                                grid.n_y, grid.win_x, grid.win_y, grid.stride_x, grid.stride_y, k0, k1, 1, _stream(dev))
    _lib.check(rc, None, "espnet_stitch_grid_band")


class PeerGather:
    """Zero-copy band placement over NVLink ("the stitch scatter kernel may write directly into a peer-mapped
    slide buffer, which IS the fused scatter + gather").  Rank 0 allocates the level-0 slide mask and a small staging area for
    the tile-overlap strips through espnet_peer_alloc and broadcasts their CUDA IPC handles; every other rank (process) maps
    them with espnet_peer_open and its stitch kernel writes its band's rows straight into rank 0's memory.  Only two tiny
    barriers and the max-merge of the strips (win - stride rows per rank boundary) remain of the "gather".  Build once per
    (slide size, tiling, process group) -- the handle exchange is a collective -- and pass to segment_slide(gather=...).
    Raises (on every rank) if the GPUs cannot peer; callers then fall back to gather_bands (NCCL send / recv)."""

    def __init__(self, grid: TileGrid, slide_h: int, slide_w: int, rank: int, world: int, device: torch.device, ws: int = 2400, group=None,
                 balance: str = "tiles"):
        import torch.distributed as dist
        self.grid, self.sh, self.sw, self.rank, self.world, self.group, self.dist = grid, slide_h, slide_w, rank, world, group, dist
        self.device, self.balance = device, balance
        self.plan = band_plan(grid, slide_h, world, balance)
        self.strip_rows = max([sp - y0 for y0, sp, _ in self.plan] + [1])
        n0, n1 = slide_h * slide_w, world * self.strip_rows * slide_w
        self._ptr0 = self._ptr1 = None
        L = _lib.lib()

        def all_ok(ok: bool, what: str):
            # failures must be collective: a rank that raised alone would leave the others waiting in the next barrier
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                raise RuntimeError("PeerGather: %s failed on at least one rank (no CUDA IPC / peer access between the GPUs?)" % what)

        dev0 = [device.index]
        dist.broadcast_object_list(dev0, src=0, group=group, device=device)         # which GPU rank 0 sits on
        all_ok(rank == 0 or (dev0[0] != device.index and torch.cuda.can_device_access_peer(device.index, dev0[0])), "the peer-access capability check")
        payload = [None, None]
        ok = True
        if rank == 0:
            p0, p1 = C.c_void_p(), C.c_void_p()
            h0, h1 = C.create_string_buffer(64), C.create_string_buffer(64)
            ok = L.espnet_peer_alloc(n0, device.index, C.byref(p0), h0) == _lib.OK and L.espnet_peer_alloc(n1, device.index, C.byref(p1), h1) == _lib.OK
            if ok:
                self._ptr0, self._ptr1 = p0.value, p1.value
                payload = [h0.raw, h1.raw]
        all_ok(ok, "allocating the exported slide mask")
        dist.broadcast_object_list(payload, src=0, group=group, device=device)
        ok = True
        if rank != 0:
            p0, p1 = C.c_void_p(), C.c_void_p()
            ok = L.espnet_peer_open(payload[0], device.index, C.byref(p0)) == _lib.OK and L.espnet_peer_open(payload[1], device.index, C.byref(p1)) == _lib.OK
            if ok:
                self._ptr0, self._ptr1 = p0.value, p1.value
        all_ok(ok, "mapping rank 0's slide mask")
        if rank == 0:       # rank 0 sees its own buffers as ordinary tensors (aliases, not owners)
            self.level0 = torch.as_tensor(_DeviceBytes(self._ptr0, n0), device=device).view(slide_h, slide_w)
            self.strips = torch.as_tensor(_DeviceBytes(self._ptr1, n1), device=device).view(world, self.strip_rows, slide_w)
        # every rank proves with the stitch kernel itself (running on ITS GPU) that it can write rank 0's memory
        probe = TileGrid(1, 1, 8, 1, 8, 1)
        tile = torch.full((1, 1, 8), rank + 1, dtype=torch.uint8, device=device)
        _stitch_raw(self._ptr1 + rank * self.strip_rows * slide_w, 0, 1, 1, slide_w, tile, probe, 0, 1, 8)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        ok = True
        if rank == 0:
            ok = self.strips[:, 0, 0].cpu().tolist() == [r + 1 for r in range(world)]
            self.strips.zero_()
            torch.cuda.synchronize(device)
        all_ok(ok, "writing through the peer mapping")

    def close(self):
        """Collective: unmap on the peers, then free on rank 0."""
        if self._ptr0 is None:
            return
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
        if self.rank != 0:
            L.espnet_peer_close(self._ptr0, self.device.index)
            L.espnet_peer_close(self._ptr1, self.device.index)
        self.dist.barrier(group=self.group)
        if self.rank == 0:
            self.level0 = self.strips = None
            L.espnet_peer_free(self._ptr0, self.device.index)
            L.espnet_peer_free(self._ptr1, self.device.index)
        self._ptr0 = self._ptr1 = None

    def place(self, tile_masks: Optional[torch.Tensor], k0: int, k1: int, ws: int) -> dict:
        """Collective: every rank stitches its tiles [k0, k1) into rank 0's slide mask (rows it owns) / strip staging (rows shared
        with an earlier band); rank 0 merges the strips.  Returns byte counts."""
        y0, split, y1 = self.plan[self.rank]
        self.dist.barrier(group=self.group)          # rank 0 is done with the previous result
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        if k1 > k0 and y1 > y0:
            if split > y0:
                _stitch_raw(self._ptr1 + self.rank * self.strip_rows * self.sw, y0, split - y0, self.sh, self.sw, tile_masks, self.grid, k0, k1, ws)
            if y1 > split:
                _stitch_raw(self._ptr0 + split * self.sw, split, y1 - split, self.sh, self.sw, tile_masks, self.grid, k0, k1, ws)
        ev[1].record()
        self.dist.barrier(group=self.group)          # every band has landed (kernel completion makes the peer writes visible)
        # `_place_events`: the stitch kernels alone (this rank's writes into rank 0's memory), without the barriers around them
        stats = {"bytes_received": 0, "bytes_merged": 0, "gather": "p2p", "_place_events": ev}
        if self.rank == 0:
            for r, (a, sp, b) in enumerate(self.plan):
                if r and sp > a:
                    max_merge_(self.level0[a:sp], self.strips[r, :sp - a])
                    stats["bytes_merged"] += (sp - a) * self.sw
                if r:
                    stats["bytes_received"] += (b - a) * self.sw
        return stats


def segment_slide(model, slide_u8: torch.Tensor, mean, std, std_size: float = 512, mpp: float = 1.0, overlap: float = 0.1,
                  downsample: float = 1.0, ws: int = 2400, batch: int = 256, rank: int = 0, world: int = 1,
                  reduce_to_rank0: bool = True, slide_y0: int = 0, slide_h: Optional[int] = None, timings: Optional[dict] = None,
                  gather: Optional["PeerGather"] = None, balance: str = "tiles"):
    """Overlapping-tile WSI segmentation (BASELINE.json config 4): T1 tiles -> ESPNet forward + arg-max per
    tile -> T3 max-merge -> T4 /8 mask.  `slide_u8` is the resident BGR slide uint8 [rows,SW,3]: the whole slide, or -- with
    `slide_h` (full height) and `slide_y0` given -- just the rows [slide_y0, slide_y0 + rows) that this rank's tiles read
    (`band_tiles`).  With world > 1 the tiles are sharded across ranks as contiguous, equally sized ranges of scan_region's
    row-major order (balance="tiles"; "rows" hands out whole tile rows); the forward has no collective; every rank stitches its
    own band, either straight into rank 0's slide mask over NVLink (`gather` = a PeerGather) or into a local band buffer that
    `gather_bands` ships to rank 0 (NCCL send / recv).
    Returns (level0 uint8 [SH,SW] on rank 0 (the rank's band mask [y1-y0,SW] or None elsewhere; the band everywhere when
    reduce_to_rank0=False), ds8 uint8 [int(SH/8), int(SW/8)] on rank 0, n_local_tiles)."""
    if downsample != 1.0:
        raise RuntimeError("only level-0 tiling (downsample 1) is wired to the resident-slide reader")
    _check_u8(slide_u8, "slide_u8", 3)
    sh = int(slide_u8.shape[0]) + slide_y0 if slide_h is None else int(slide_h)
    sw = int(slide_u8.shape[1])
    dev = slide_u8.device
    grid = tile_grid(sw, sh, std_size, mpp, mpp, overlap, downsample)
    if grid.win_x % 8 or grid.win_y % 8:
        raise RuntimeError("tile size %dx%d is not a multiple of 8 (ESPNet needs it, Model.py cat at :373)" % (grid.win_x, grid.win_y))
    k0, k1, y0, y1 = band_tiles(grid, sh, rank, world, balance)
    n_local = k1 - k0
    if n_local and (slide_y0 > y0 or slide_y0 + int(slide_u8.shape[0]) < y1):
        raise RuntimeError("slide_u8 holds rows [%d,%d) but this rank's tiles read rows [%d,%d)" % (slide_y0, slide_y0 + slide_u8.shape[0], y0, y1))
    p2p = gather is not None and world > 1 and reduce_to_rank0
    if p2p and (gather.grid != grid or gather.sh != sh or gather.sw != sw or gather.world != world or gather.balance != balance):
        raise RuntimeError("the PeerGather was built for another slide / tiling / world size / balance")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
    if ev:
        ev[0].record()
    masks = None
    if n_local:
        org = grid.origins_range(k0, k1)
        org[:, 1] -= slide_y0                              # tile origins relative to the resident rows
        origins = torch.from_numpy(org).to(dev)
        masks = torch.empty((n_local, grid.win_y, grid.win_x), dtype=torch.uint8, device=dev)
        # rows below the resident part are outside the SLIDE for the last band only, where zero padding is what read_region does
        for s in range(0, n_local, batch):
            e = min(s + batch, n_local)
            model.segment_tiles(slide_u8, origins[s:e], grid.win_y, grid.win_x, mean, std, out=masks[s:e])
    if ev:
        ev[1].record()
    band = None
    if not p2p:
        band = torch.zeros((y1 - y0, sw), dtype=torch.uint8, device=dev)
        if n_local:
            stitch_grid(band, masks, grid, 0, 0, ws, band_y0=y0, slide_h=sh, tiles=(k0, k1))
    if ev:
        ev[2].record()
    if p2p:
        stats = gather.place(masks, k0, k1, ws)
        level0 = gather.level0 if rank == 0 else None
    elif world > 1 and reduce_to_rank0:
        level0, stats = gather_bands(band, grid, sh, sw, rank, world, balance=balance)
        stats["gather"] = "nccl send/recv"
    else:
        level0, stats = band, {"bytes_received": 0, "bytes_merged": 0, "gather": "none"}
        if world == 1 and (y0 != 0 or y1 != sh):           # tiles do not reach the last slide rows (cannot happen with T1's ceil)
            level0 = torch.zeros((sh, sw), dtype=torch.uint8, device=dev)
            level0[y0:y1].copy_(band)
    if ev:
        ev[3].record()
    ds8 = downsample8(level0, ws) if (level0 is not None and tuple(level0.shape) == (sh, sw)) else None
    pe = stats.pop("_place_events", None)
    if ev:
        torch.cuda.synchronize(dev)
        if pe is not None:
            stats["place_kernels_ms"] = pe[0].elapsed_time(pe[1])
        timings.update(forward_ms=ev[0].elapsed_time(ev[1]), stitch_ms=ev[1].elapsed_time(ev[2]), gather_ms=ev[2].elapsed_time(ev[3]), **stats)
    return (level0 if level0 is not None else band), ds8, n_local
