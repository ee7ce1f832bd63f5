#!/usr/bin/env python
"""Host-link probe for the e2e metric (run under torchrun with N ranks, or alone): every rank copies the bench's per-step
payload (50.3 MB of u8 crops H2D, 16.8 MB of masks D2H) from / to pinned memory, all ranks at the same time, and reports
GB/s per rank.  Tells whether the end-to-end loss at 8 GPUs is the host link / host memory or the pipeline."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ctx = bench.Ctx()
    h_in = torch.empty(64 * 512 * 512 * 3, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(64 * 512 * 512, dtype=torch.uint8).pin_memory()
    d_in = torch.empty_like(h_in, device=ctx.dev)
    d_out = torch.empty_like(h_out, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(ctx.dev), torch.cuda.Stream(ctx.dev)
    res = {}

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream(ctx.dev)
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    for name, fn, nbytes in (("h2d", h2d, h_in.numel()), ("d2h", d2h, h_out.numel()), ("both", both, h_in.numel() + h_out.numel())):
        for _ in range(3):
            fn()
        ms = ctx.timed(fn, 20)
        res[name + "_gbs_per_rank"] = nbytes * 20 / (ms * 1e-3) / 1e9
        res[name + "_ms"] = ms / 20
    if ctx.rank == 0:
        print(json.dumps({"n_ranks": ctx.world, "host": ctx.host, "cpu_count": os.cpu_count(), **res}))
    ctx.close()


if __name__ == "__main__":
    main()
