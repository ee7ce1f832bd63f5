#!/usr/bin/env python
"""Per-kernel CUDA-event times of ONE 512 x 512 crop through the full ESPNet (batch 1, fp32-equivalent mode): where the 0.27 ms
of the batch-1 forward go.  Events serialise the kernels (no programmatic overlap), so the sum is above the chained forward."""
import json
import os
import sys

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
if os.environ.get("DEC_IMPL"):
    m.set_option("dec_impl", int(os.environ["DEC_IMPL"]))
out = {}
for B in (1, 4):
    u8 = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).to(dev)
    for _ in range(20):
        m.segment(u8, mean, std)
    torch.cuda.synchronize()
    m.profile(True)
    n = 50
    for _ in range(n):
        m.segment(u8, mean, std)
    torch.cuda.synchronize()
    rep = m.profile_report()
    m.profile(False)
    out["batch_%d" % B] = {k: {"us_per_launch": round(1e3 * v[0] / v[1], 2), "launches": v[1] // n} for k, v in rep.items()}
    out["batch_%d" % B]["sum_us"] = round(1e3 * sum(v[0] for v in rep.values()) / n, 1)
print(json.dumps(out))
