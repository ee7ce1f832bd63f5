#!/usr/bin/env python
"""Where the end-to-end step (pinned host crops -> masks on the host) differs from the device-resident forward.
ESPNet-C, 64 crops of 512 x 512, fp32-equivalent mode, the HostPipeline of Model.py (3 streams, depth 3), 40 steps each:
  resident_f32   model(x): normalised fp32 batch in HBM, logits out (bench.py `value`)
  segment_u8     model.segment(u8 in HBM): fused normalise, forward, x8 up-sample + arg-max, mask in HBM
  pipe_nocopy    the pipeline with both copies removed (events and stream hops only)
  pipe_h2d       + the host-to-device copy of every batch
  pipe_d2h       + the device-to-host copy of every mask (no H2D)
  pipe_full      both copies (bench.py `e2e`)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glomeruli_segmentation_b200 import ESPNet_Encoder, FOLD_MEAN_STD  # noqa: E402

dev = torch.device("cuda:0")
B, H, W, STEPS = 64, 512, 512, 40
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_fold1.npz"))
mean, std = FOLD_MEAN_STD[1]
m = ESPNet_Encoder(5, 2, 8)
m.load_state_dict({k[len("encoder."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("encoder.")}, strict=True)
m = m.to(dev).eval()
u8_host = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B, H, W, 3), dtype=np.uint8)).pin_memory()
mask_host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
u8 = u8_host.to(dev)
x = (((u8.float() - torch.tensor(mean, device=dev)) / torch.tensor(std, device=dev)) / 255.0).permute(0, 3, 1, 2).contiguous()
mask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)


def timed(fn, pipe=None):
    for _ in range(5):
        fn()
    if pipe is not None:
        pipe.drain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record()
    if pipe is not None:
        for s in (pipe.s_in, pipe.s_run, pipe.s_out):
            s.wait_stream(cur)
    for _ in range(STEPS):
        fn()
    if pipe is not None:
        for s in (pipe.s_in, pipe.s_run, pipe.s_out):
            cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / STEPS, 4)


def pipe_variant(h2d, d2h):
    pipe = m.host_pipeline(B, H, W, mean, std, depth=3)
    for k in range(3):
        pipe.d_in[k].copy_(u8)

    def submit():
        k = pipe.n % pipe.depth
        if pipe.n >= pipe.depth:
            pipe.s_in.wait_event(pipe.ev_run[k])
            pipe.s_run.wait_event(pipe.ev_out[k])
        with torch.cuda.stream(pipe.s_in):
            if h2d:
                pipe.d_in[k].copy_(u8_host, non_blocking=True)
            pipe.ev_in[k].record(pipe.s_in)
        with torch.cuda.stream(pipe.s_run):
            pipe.s_run.wait_event(pipe.ev_in[k])
            m.segment(pipe.d_in[k], mean, std, out=pipe.d_mask[k])
            pipe.ev_run[k].record(pipe.s_run)
        with torch.cuda.stream(pipe.s_out):
            pipe.s_out.wait_event(pipe.ev_run[k])
            if d2h:
                mask_host.copy_(pipe.d_mask[k], non_blocking=True)
            pipe.ev_out[k].record(pipe.s_out)
        pipe.n += 1
    return submit, pipe


out = {}
for rep in range(2):
    out.setdefault("resident_f32", []).append(timed(lambda: m(x)))
    out.setdefault("segment_u8", []).append(timed(lambda: m.segment(u8, mean, std, out=mask)))
    for name, (a, b) in (("pipe_nocopy", (0, 0)), ("pipe_h2d", (1, 0)), ("pipe_d2h", (0, 1)), ("pipe_full", (1, 1))):
        fn, pipe = pipe_variant(a, b)
        out.setdefault(name, []).append(timed(fn, pipe))
        del pipe
print(json.dumps({"ms_per_step": out, "batch": B, "steps": STEPS}))
