#!/usr/bin/env python
"""Multi-GPU check of the WSI path (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/multi_gpu_wsi_check.py
Every rank builds only its band of a seeded slide, segments its tile rows and stitches its band; the bands reach rank 0 either
through gather_bands (NCCL send / recv) or through a PeerGather (the stitch kernels write into rank 0's slide mask over NVLink).
Rank 0 also computes the WHOLE slide alone and demands bit-equal level-0 and /8 masks from both paths."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from glomeruli_segmentation_b200 import FOLD_MEAN_STD, wsi  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = type("C", (), {"dev": dev})()
    model = bench.make_model(ctx, False, 1, "fp32")
    mean, std = FOLD_MEAN_STD[1]
    out = []
    for (sw, sh, ov) in ((4096, 3072, 0.1), (3000, 5000, 0.5), (1500, 1100, 0.25)):
        grid = wsi.tile_grid(sw, sh, 512, 1.0, 1.0, ov, 1.0)
        k0, k1, y0, y1 = wsi.band_tiles(grid, sh, rank, world)
        band_slide = bench.synth_slide_rows(dev, sw, y0, y1, seed=7, block=512)
        whole = bench.synth_slide_rows(dev, sw, 0, sh, seed=7, block=512) if rank == 0 else None
        ref = wsi.segment_slide(model, whole, mean, std, overlap=ov, batch=64) if rank == 0 else None
        try:                                             # collective: raises on every rank or on none
            pg = wsi.PeerGather(grid, sh, sw, rank, world, dev)
        except RuntimeError as e:
            pg = None
            if rank == 0:
                out.append({"slide": [sw, sh], "mode": "p2p", "unavailable": str(e)})
        for mode in ("nccl", "p2p", "p2p"):              # the peer-mapped buffers are reused across steps: run that path twice
            if mode == "p2p" and pg is None:
                continue
            tm = {}
            level0, ds8, n_local = wsi.segment_slide(model, band_slide, mean, std, overlap=ov, batch=64, rank=rank, world=world,
                                                     slide_y0=y0, slide_h=sh, timings=tm, gather=pg if mode == "p2p" else None)
            n = torch.tensor([n_local], device=dev)
            dist.all_reduce(n)
            if rank == 0:
                ref0, ref8, n_all = ref
                ok = bool(torch.equal(level0, ref0)) and bool(torch.equal(ds8, ref8)) and int(n.item()) == n_all == grid.count
                out.append({"slide": [sw, sh], "overlap": ov, "tiles": grid.count, "mode": mode, "equal_to_single_gpu": ok, "gather": tm})
                assert ok, out[-1]
        level0 = ds8 = None
        if pg is not None:
            pg.close()
    if rank == 0:
        print(json.dumps({"world": world, "cases": out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
