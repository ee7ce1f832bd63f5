#!/usr/bin/env python
"""Minimal driver for ncu captures: a few forwards of the bench workload (ESPNet-C, 512x512) in one compute mode.
usage: python profiles/run_forward.py [fp32|f16tc] [batch] [iters]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glomeruli_segmentation_b200 import ESPNet_Encoder, FOLD_MEAN_STD  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_fold1.npz"))
sd = {k[len("encoder."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("encoder.")}
m = ESPNet_Encoder(5, 2, 8)
m.load_state_dict(sd, strict=True)
m = m.to("cuda:0").eval().set_mode(mode)
mean, std = FOLD_MEAN_STD[1]
u8 = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).to("cuda:0")
x = (((u8.float() - torch.tensor(mean, device="cuda:0")) / torch.tensor(std, device="cuda:0")) / 255.0).permute(0, 3, 1, 2).contiguous()
for _ in range(iters):
    y = m(x)
torch.cuda.synchronize()
print("ok", mode, B, tuple(y.shape), float(y.abs().max()))
