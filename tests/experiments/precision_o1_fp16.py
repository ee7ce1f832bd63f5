#!/usr/bin/env python
"""Precision experiment behind profiles/r02_level3_block_budget.md (runs on the CPU oracle, test infrastructure):
how large does the logit error get when ONLY the reduced map o1 of the ESP / DownSampler blocks is rounded to fp16
(which would remove the A_lo term of the 3-term split), everything else fp32?

  python tests/experiments/precision_o1_fp16.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_fold_state_dict  # noqa: E402
from oracle import espnet_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
MODE = {"o1": None}
_branches = O._branches


def rnd16(x, scale=0.25):
    return ((x * scale).half().float()) / scale


def patched(sd, key, o1):
    if MODE["o1"] == "all" or (MODE["o1"] == "l3" and "level3" in key):
        o1 = rnd16(o1)
    return _branches(sd, key, o1)


O._branches = patched
for fold in (1, 3, 5):
    sd = load_fold_state_dict(fold)
    mean, std = O.FOLD_MEAN_STD[fold]
    for dist, sig in (("D1", 0), ("D2", 2.0), ("D2", 16.0)):
        u8 = O.synth_crops(dist, 2, 256, 256, seed=5, sigma=sig or 4.0)
        x = torch.from_numpy(O.normalise_bgr_u8(u8, mean, std))
        MODE["o1"] = None
        ref = O.espnet_forward(sd, x)
        row = {}
        for m in ("l3", "all"):
            MODE["o1"] = m
            y = O.espnet_forward(sd, x)
            row[m] = ((y - ref).abs().max().item(), (y.argmax(1) == ref.argmax(1)).float().mean().item())
        print("fold %d %s sigma %4.1f  o1->fp16 at level 3: max-abs %.2e agree %.5f | everywhere: %.2e %.5f"
              % (fold, dist, sig, row["l3"][0], row["l3"][1], row["all"][0], row["all"][1]))
