#!/usr/bin/env python
"""Host-side cost of enqueueing one forward (full ESPNet, 16 crops of 512 x 512): wall clock of 20 forwards enqueued without
a synchronise (600 launches, below the driver's launch-queue depth, while the GPU needs ~0.9 ms per forward) / 20.
Printed with and without the per-forward stage table ("stages" option) and through the Python wrapper vs graph replay."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glomeruli_segmentation_b200 import ESPNet, FOLD_MEAN_STD  # noqa: E402

dev = "cuda:0"
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_fold1.npz"))
mean, std = FOLD_MEAN_STD[1]
m = ESPNet(5, 2, 8)
m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
m = m.to(dev).eval()
u8 = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (16, 512, 512, 3), dtype=np.uint8)).to(dev)
out = torch.empty((16, 512, 512), dtype=torch.uint8, device=dev)


def enqueue_us(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        best = min(best, (time.perf_counter() - t0) / n * 1e6)
        torch.cuda.synchronize()
    return round(best, 1)


rec = {"python_segment_us": enqueue_us(lambda: m.segment(u8, mean, std, out=out))}
g = m.capture(16, 512, 512, mean, std)
g.input.copy_(u8)
rec["graph_run_us"] = enqueue_us(g.run)
print(json.dumps(rec))
