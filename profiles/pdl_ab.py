#!/usr/bin/env python
"""A/B of the "pdl" option (programmatic dependent launch between the kernels of a forward), per kernel class.
usage: python profiles/pdl_ab.py [mask ...]      masks: bit 1 stem, 2 unused (was pool_b2, now part of the stem), 4 3x3-s2 reduce, 8 branch, 16 1x1 reduce, 32 head / decoder,
                                                 64 last kernel of the forward; PDL_AB_FAST=1: batch 64 only
Prints one JSON record: per mask the CUDA-event time of the ESPNet-C and full ESPNet forwards at batch 64 (fp32-equivalent mode,
inputs resident, 30 forwards back to back) and of one full forward at batch 1, plain launches and CUDA-graph replay."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glomeruli_segmentation_b200 import ESPNet, ESPNet_Encoder, FOLD_MEAN_STD  # noqa: E402

masks = [int(a) for a in sys.argv[1:]] or [0, 127]
FAST = os.environ.get("PDL_AB_FAST") == "1"
dev = "cuda:0"
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_fold1.npz"))
mean, std = FOLD_MEAN_STD[1]
enc = ESPNet_Encoder(5, 2, 8)
enc.load_state_dict({k[len("encoder."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("encoder.")}, strict=True)
enc = enc.to(dev).eval()
full = ESPNet(5, 2, 8)
full.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
full = full.to(dev).eval()
rng = np.random.default_rng(0)
u8 = torch.from_numpy(rng.integers(0, 256, (64, 512, 512, 3), dtype=np.uint8)).to(dev)
x = (((u8.float() - torch.tensor(mean, device=dev)) / torch.tensor(std, device=dev)) / 255.0).permute(0, 3, 1, 2).contiguous()
lg = torch.empty((64, 5, 512, 512), device=dev)


def timed(fn, n, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
ref_mask = None
for rep in range(2):
    for m in masks:
        enc.set_option("pdl", m)
        full.set_option("pdl", m)
        r = out.setdefault(str(m), {"enc_b64_ms": [], "full_b64_ms": [], "b1_plain_ms": [], "b1_graph_ms": []})
        r["enc_b64_ms"].append(round(timed(lambda: enc(x), 30), 4))
        r["full_b64_ms"].append(round(timed(lambda: full.segment(u8, mean, std, logits=lg), 30), 4))
        mk = full.segment(u8, mean, std, logits=lg).clone()
        if ref_mask is None:
            ref_mask, ref_lg = mk, lg.clone()
        r["identical_to_first"] = bool(torch.equal(mk, ref_mask) and torch.equal(lg, ref_lg))
        if FAST:
            continue
        r["b1_plain_ms"].append(round(timed(lambda: full.segment(u8[:1], mean, std), 200, 20), 4))
        g = full.capture(1, 512, 512, mean, std)
        g.input.copy_(u8[:1])
        r["b1_graph_ms"].append(round(timed(g.run, 200, 20), 4))
        del g
print(json.dumps(out))
if FAST:
    sys.exit(0)

# second record: where the option starts to cost -- full forward, plain launches, batch sweep, all classes on / off
sweep = {}
for b in (1, 2, 4, 8, 16, 32, 64):
    row = {}
    for m in (0, 127, 56):
        full.set_option("pdl", m)
        row[str(m)] = round(min(timed(lambda: full.segment(u8[:b], mean, std), 100 if b <= 8 else 30, 10) for _ in range(2)), 4)
    sweep[str(b)] = row
print(json.dumps({"full_forward_ms_by_batch": sweep}))
