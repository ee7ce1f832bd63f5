#!/usr/bin/env python
"""Minimal driver for ncu captures: a few forwards of the bench workloads (512x512 crops) in one compute mode.
usage: python profiles/run_forward.py [fp32|f16tc] [batch] [iters] [encoder|full|aux]
  encoder  ESPNet-C forward (the headline workload) + x8 up-sample / arg-max
  full     full ESPNet: u8 crops in (fused normalise), logits + arg-max out
  aux      the kernels either side of the forward at realistic sizes: slide tile reader, grid / box stitch, /8 LUT gather, render,
           confusion histogram, class counts, crop front-end"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glomeruli_segmentation_b200 import ESPNet, ESPNet_Encoder, FOLD_MEAN_STD, frontend, iouEval, wsi  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
net = sys.argv[4] if len(sys.argv) > 4 else "encoder"
dev = "cuda:0"
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_fold1.npz"))
mean, std = FOLD_MEAN_STD[1]
u8 = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).to(dev)
if net == "encoder":
    sd = {k[len("encoder."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("encoder.")}
    m = ESPNet_Encoder(5, 2, 8)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval().set_mode(mode)
    x = (((u8.float() - torch.tensor(mean, device=dev)) / torch.tensor(std, device=dev)) / 255.0).permute(0, 3, 1, 2).contiguous()
    y = None
    for _ in range(iters):
        y = m(x)
    mask = m.segment(u8, mean, std)
    torch.cuda.synchronize()
    print("ok", net, mode, B, None if y is None else (tuple(y.shape), float(y.abs().max())), int(mask.sum()))
else:
    m = ESPNet(5, 2, 8)
    m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
    m = m.to(dev).eval().set_mode(mode)
    lg = torch.empty((B, 5, 512, 512), device=dev)
    for _ in range(iters):
        mask = m.segment(u8, mean, std, logits=lg)
    torch.cuda.synchronize()
    print("ok", net, mode, B, float(lg.abs().max()), int(mask.sum()))
    if net == "aux":
        sw, sh = 16000, 12000
        slide = torch.randint(0, 256, (sh, sw, 3), dtype=torch.uint8, device=dev)
        level0, ds8, n = wsi.segment_slide(m, slide, mean, std, batch=64)                      # slide reader stem, stitch_grid, /8 LUT gather
        img = frontend.render_slide_ds8(slide, level0)                                          # render_ds8
        color, over = frontend.colorize_overlay(mask[0], u8[0])                                 # palette_overlay
        ev = iouEval(5)
        ev.addBatch(mask, torch.roll(mask, 1, 0))                                               # confusion_hist
        counts = frontend.class_pixel_counts(mask)                                              # class_count
        crops = torch.randint(0, 256, (8, 700, 900, 3), dtype=torch.uint8, device=dev)
        xin = frontend.preprocess_resize(crops, mean, std, 1024, 512)                           # preprocess_resize
        back = frontend.resize_mask_nearest(mask[:8], 700, 900)                                 # resize_nearest_u8
        boxes = [[100 * i, 80 * i, 100 * i + 600 + 10 * i, 80 * i + 500, 1.0] for i in range(8)]
        xb = frontend.preprocess_boxes(slide, boxes, mean, std, 1024, 512)                      # preprocess_resize_boxes
        full = torch.zeros((sh, sw), dtype=torch.uint8, device=dev)
        wsi.stitch_boxes(full, boxes, [torch.randint(0, 5, (int(b[3] - b[1]), int(b[2] - b[0])), dtype=torch.uint8, device=dev) for b in boxes])
        wsi.max_merge_(full[:52], level0[:52].contiguous())                                     # max_merge_u8
        torch.cuda.synchronize()
        print("ok aux", n, tuple(img.shape), counts.sum().item(), tuple(xin.shape), tuple(xb.shape), tuple(back.shape))
