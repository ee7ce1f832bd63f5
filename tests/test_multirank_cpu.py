"""N > 1 path on CPU (gloo, world_size 2): the tile rows of a slide shard across ranks with no collective in the
forward; the only exchange is the final max-reduce of the level-0 masks to rank 0 (wsi.segment_slide).
The per-tile class maps are synthetic here (the forward itself needs the GPU); what is checked is the host logic a
multi-GPU run relies on: shard_rows partitions the tile rows, TileGrid.origins enumerates the same tiles as the
reference's scan_region loop, and max-merging per-rank band masks + MAX-reduce equals the single-process stitch."""
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


def _worker(rank, world, port, sw, sh, ov, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = wsi.tile_grid(sw, sh, 64, 1.0, 1.0, ov, 1.0)
        row0, rows = wsi.shard_rows(grid.n_y, rank, world)
        ylim = wsi.stitch_y_limit(sw, sh, 2400)
        local = torch.from_numpy(_band_mask(sw, sh, grid, row0, rows, ylim))
        dist.reduce(local, dst=0, op=dist.ReduceOp.MAX)          # the one exchange of the WSI path
        counts = torch.tensor([rows * grid.n_x], dtype=torch.int64)
        dist.all_reduce(counts)
        if rank == 0:
            q.put((local.numpy(), int(counts.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sw,sh,ov", [(500, 380, 0.1), (333, 420, 0.5)])
def test_two_rank_band_sharding_equals_single_process(sw, sh, ov):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sw, sh, ov, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, n_tiles = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    grid = wsi.tile_grid(sw, sh, 64, 1.0, 1.0, ov, 1.0)
    # the sharded tile set is the reference's tile set (detect_glomus_test.py:264-304 restated in the oracle)
    origins, n_x, n_y, win_x, win_y, stride_x, stride_y = W.tile_grid(sw, sh, 64, 1.0, 1.0, ov, 1.0)
    assert (grid.n_x, grid.n_y, grid.win_x, grid.win_y, grid.stride_x, grid.stride_y) == (n_x, n_y, win_x, win_y, stride_x, stride_y)
    bands = [wsi.shard_rows(grid.n_y, r, 2) for r in range(2)]
    assert np.array_equal(np.concatenate([grid.origins(r0, n) for r0, n in bands]), origins)
    assert n_tiles == grid.count
    single = _band_mask(sw, sh, grid, 0, grid.n_y, wsi.stitch_y_limit(sw, sh, 2400))
    assert np.array_equal(merged, single)
