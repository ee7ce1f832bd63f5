"""N > 1 path on CPU (gloo, world_size 2 and 3): the tile rows of a slide shard across ranks with no collective in the
forward; every rank stitches only its own band of slide rows and the only exchange is wsi.gather_bands:
point-to-point, rows no earlier band covers are received straight into the slide mask, the tile-overlap rows are
max-merged.  The per-tile class maps are synthetic here (the forward itself needs the GPU); what is checked is the host
logic a multi-GPU run relies on: shard_rows / band_rows partition the slide, TileGrid.origins enumerates the same tiles as
the reference's scan_region loop, and the gathered mask equals the single-process stitch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from glomeruli_segmentation_b200 import wsi
from oracle import wsi_oracle as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _tile_mask(k, win_y, win_x):
    """Deterministic synthetic class map of tile k (values 0..4)."""
    rng = np.random.default_rng(1000 + k)
    coarse = rng.integers(0, 5, (win_y // 16 + 1, win_x // 16 + 1), dtype=np.uint8)
    return np.kron(coarse, np.ones((16, 16), np.uint8))[:win_y, :win_x]


def _band_mask(sw, sh, grid, row0, rows, y_limit):
    """CPU restatement of the stitch for a band of tile rows: slide[y,x] = max over covering tiles (T3)."""
    out = np.zeros((sh, sw), np.uint8)
    org = grid.origins(row0, rows)
    for idx, (x0, y0) in enumerate(org):
        k = (row0 + idx // grid.n_x) * grid.n_x + idx % grid.n_x
        m = _tile_mask(k, grid.win_y, grid.win_x)
        x1, y1 = min(x0 + grid.win_x, sw), min(min(y0 + grid.win_y, sh), y_limit)
        if x1 <= x0 or y1 <= y0:
            continue
        out[y0:y1, x0:x1] = np.maximum(out[y0:y1, x0:x1], m[: y1 - y0, : x1 - x0])
    return out


def _range_mask(sw, sh, grid, k0, k1, y_limit):
    """The same for a row-major tile index range [k0, k1) (balance="tiles")."""
    out = np.zeros((sh, sw), np.uint8)
    for k, (x0, y0) in zip(range(k0, k1), grid.origins_range(k0, k1)):
        m = _tile_mask(k, grid.win_y, grid.win_x)
        x1, y1 = min(x0 + grid.win_x, sw), min(min(y0 + grid.win_y, sh), y_limit)
        if x1 > x0 and y1 > y0:
            out[y0:y1, x0:x1] = np.maximum(out[y0:y1, x0:x1], m[: y1 - y0, : x1 - x0])
    return out


def _worker(rank, world, port, sw, sh, ov, q, balance="rows"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = wsi.tile_grid(sw, sh, 64, 1.0, 1.0, ov, 1.0)
        k0, k1, y0, y1 = wsi.band_tiles(grid, sh, rank, world, balance)
        ylim = wsi.stitch_y_limit(sw, sh, 2400)
        # the rank's band buffer holds ONLY slide rows [y0, y1)
        band = torch.from_numpy(_range_mask(sw, sh, grid, k0, k1, ylim)[y0:y1].copy())
        level0, stats = wsi.gather_bands(band, grid, sh, sw, rank, world, merge=lambda d, s: d.copy_(torch.maximum(d, s)), balance=balance)
        counts = torch.tensor([k1 - k0], dtype=torch.int64)
        dist.all_reduce(counts)
        if rank == 0:
            q.put((level0.numpy(), int(counts.item()), stats))
        else:
            assert level0 is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sw,sh,ov,world,balance", [(500, 380, 0.1, 2, "rows"), (333, 420, 0.5, 2, "tiles"), (300, 700, 0.5, 3, "tiles"),
                                                    (200, 130, 0.1, 3, "rows"), (410, 300, 0.1, 3, "tiles")])
def test_band_sharding_and_gather_equal_single_process(sw, sh, ov, world, balance):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sw, sh, ov, q, balance)) for r in range(world)]
    for p in procs:
        p.start()
    merged, n_tiles, stats = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    grid = wsi.tile_grid(sw, sh, 64, 1.0, 1.0, ov, 1.0)
    # the sharded tile set is the reference's tile set (detect_glomus_test.py:264-304 restated in the oracle)
    origins, n_x, n_y, win_x, win_y, stride_x, stride_y = W.tile_grid(sw, sh, 64, 1.0, 1.0, ov, 1.0)
    assert (grid.n_x, grid.n_y, grid.win_x, grid.win_y, grid.stride_x, grid.stride_y) == (n_x, n_y, win_x, win_y, stride_x, stride_y)
    ranges = [wsi.band_tiles(grid, sh, r, world, balance)[:2] for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == grid.count and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert np.array_equal(np.concatenate([grid.origins_range(a, b) for a, b in ranges]), origins)
    if balance == "tiles":
        assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1          # an even split of the tiles
    assert n_tiles == grid.count
    single = _band_mask(sw, sh, grid, 0, grid.n_y, wsi.stitch_y_limit(sw, sh, 2400))
    assert np.array_equal(merged, single)
    # only the overlap rows went through the merge, and nobody shipped a whole-slide mask
    geo = [wsi.band_tiles(grid, sh, r, world, balance) for r in range(world)]
    assert stats["bytes_received"] == sum((y1 - y0) * sw for k0, k1, y0, y1 in geo[1:] if k1 > k0)
    assert stats["bytes_merged"] < stats["bytes_received"] or world == 1


def test_band_rows_cover_the_slide_and_overlap_by_the_tile_overlap():
    grid = wsi.tile_grid(40000, 30000, 512, 1.0, 1.0, 0.1, 1.0)
    geo = [wsi.band_rows(grid, 30000, r, 8) for r in range(8)]
    assert geo[0][2] == 0 and geo[-1][3] == 30000
    for a, b in zip(geo, geo[1:]):
        assert a[3] - b[2] == grid.win_y - grid.stride_y == 52          # adjacent bands share exactly the tile overlap
    assert sum(g[1] for g in geo) == grid.n_y
