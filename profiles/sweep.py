#!/usr/bin/env python
"""BASELINE configs[4]: throughput sweep crop size {256, 512, 1024} x batch {1 .. 256} on one GPU, both compute modes,
next to the CPU oracle on the box's host cores (bounded: 512x512, batch 8).  CUDA-event timing, inputs resident in HBM,
>= 3 warm-up passes.  usage: python profiles/sweep.py > profiles/rNN_sweep.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glomeruli_segmentation_b200 import ESPNet, FOLD_MEAN_STD  # noqa: E402

dev = torch.device("cuda", 0)
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_fold1.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
mean, std = FOLD_MEAN_STD[1]
model = ESPNet(5, 2, 8)
model.load_state_dict(sd, strict=True)
model = model.to(dev).eval()
rows = []
for mode in ("fp32", "f16tc"):
    model.set_mode(mode)
    for size in (256, 512, 1024):
        for B in (1, 4, 16, 64, 256):
            if size == 1024 and B > 64:
                continue
            u8 = torch.randint(0, 256, (B, size, size, 3), dtype=torch.uint8, device=dev)
            out = torch.empty((B, size, size), dtype=torch.uint8, device=dev)
            for _ in range(3):
                model.segment(u8, mean, std, out=out)
            iters = max(3, min(50, int(2000 // max(1, B * (size // 256) ** 2))))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                model.segment(u8, mean, std, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            rows.append({"mode": mode, "crop": size, "batch": B, "ms_per_batch": ms, "crops_per_s": B / ms * 1e3,
                         "mpx_per_s": B * size * size / ms * 1e3 / 1e6})
            del u8, out
cpu = None
if "--no-cpu" not in sys.argv:
    from oracle import espnet_oracle as O   # CPU baseline leg only
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.from_numpy(O.normalise_bgr_u8(np.random.default_rng(0).integers(0, 256, (8, 512, 512, 3), dtype=np.uint8), mean, std))
    O.espnet_forward(sd, x)
    t0 = time.perf_counter()
    for _ in range(2):
        O.argmax_mask(O.espnet_forward(sd, x))
    dt = (time.perf_counter() - t0) / 2
    cpu = {"crop": 512, "batch": 8, "ms_per_batch": dt * 1e3, "crops_per_s": 8 / dt, "cores": torch.get_num_threads(), "kind": "port"}
print(json.dumps({"workload": "full ESPNet(5,2,8) fold1, u8 crops resident in HBM -> u8 class maps (normalise + forward + arg-max)",
                  "gpu": torch.cuda.get_device_name(0), "rows": rows, "cpu_baseline": cpu}, indent=1))
