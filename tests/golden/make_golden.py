#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REAL REFERENCE
(`/root/reference/module/espnet/test/Model.py` + `models/espnet_fold*.pth`) in the build container.

The reference is Python and cannot travel to the GPU box, so its outputs on seeded inputs are
committed as small fixtures, together with this script.  Run:  python tests/golden/make_golden.py

Writes
  weights_fold{1..5}.npz     the shipped state_dicts, key-for-key (fp32; num_batches_tracked int64)
  espnet_golden.npz          inputs (u8 BGR crops), reference logits / masks / per-stage hook taps
  iou_golden.npz             reference iouEval histogram + getMetricRight on seeded label maps
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "module/espnet/test"))
sys.path.insert(0, os.path.join(REF, "module/common"))
sys.path.insert(0, ROOT)

import Model as RefModel  # noqa: E402  (the reference network, unmodified)
from IOUEval import iouEval  # noqa: E402
from oracle import espnet_oracle as O  # noqa: E402  (only for the seeded input generators / P0)

TAPS = ["encoder.level1", "encoder.b1", "encoder.level2_0", "encoder.level2.0", "encoder.level2.1",
        "encoder.b2", "encoder.level3_0", "encoder.level3.0", "encoder.level3.7", "encoder.b3",
        "encoder.classifier", "up_l3", "level3_C", "combine_l2_l3", "up_l2", "conv"]


def load_ref(fold):
    sd = torch.load(os.path.join(REF, "models/espnet_fold%d.pth" % fold), map_location="cpu", weights_only=True)
    m = RefModel.ESPNet(classes=5, p=2, q=8)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m, sd


def ref_preprocess(img_u8, mean, std):
    """Literal VisualizeResults_iou.py:107-119 for one crop of network size."""
    img = img_u8.astype(np.float32)
    for j in range(3):
        img[:, :, j] -= mean[j]
    for j in range(3):
        img[:, :, j] /= std[j]
    img /= 255
    return img.transpose((2, 0, 1))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {}
    # ---- weights -------------------------------------------------------------------------
    models = {}
    for k in range(1, 6):
        m, sd = load_ref(k)
        models[k] = m
        np.savez(os.path.join(HERE, "weights_fold%d.npz" % k), **{n: v.numpy() for n, v in sd.items()})

    # ---- small crops, folds 1 and 3, D1 + D2 ----------------------------------------------
    small = np.concatenate([O.synth_crops("D1", 1, 64, 96, seed=11),
                            O.synth_crops("D2", 1, 64, 96, seed=12, sigma=2.0),
                            O.synth_crops("D3", 1, 64, 96, seed=13)], 0)
    out["small_u8"] = small
    for k in (1, 3):
        mean, std = O.FOLD_MEAN_STD[k]
        x = torch.from_numpy(np.stack([ref_preprocess(c, mean, std) for c in small]))
        out["small_x_fold%d" % k] = x.numpy()
        with torch.no_grad():
            y = models[k](x)
        out["small_logits_fold%d" % k] = y.numpy()
        out["small_mask_fold%d" % k] = np.stack([y[i].max(0)[1].byte().numpy() for i in range(len(small))])

    # ---- one mid-size crop (level-3 map 24x32 so the d=8/16 taps land in-bounds), fold 1, with taps
    mid = O.synth_crops("D2", 1, 192, 256, seed=21, sigma=4.0)
    out["mid_u8"] = mid
    mean, std = O.FOLD_MEAN_STD[1]
    x = torch.from_numpy(np.stack([ref_preprocess(c, mean, std) for c in mid]))
    taps = {}
    hooks = []
    named = dict(models[1].named_modules())
    for name in TAPS:
        hooks.append(named[name].register_forward_hook(lambda mod, i, o, name=name: taps.__setitem__(name, o.detach().numpy())))
    with torch.no_grad():
        y = models[1](x)
    for h in hooks:
        h.remove()
    out["mid_logits_fold1"] = y.numpy()
    out["mid_mask_fold1"] = y[0].max(0)[1].byte().numpy()[None]
    for name, v in taps.items():
        # big maps are kept as fp16 (stage localisation only); the small ones stay fp32 (pins)
        out["mid_tap:" + name] = v if v.size <= 100_000 else v.astype(np.float16)
    # ESPNet-C (encoder-only, modelType 2): encoder.* weights + x8 bilinear upsample + argmax
    enc = RefModel.ESPNet_Encoder(classes=5, p=2, q=8)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in models[1].state_dict().items() if k.startswith("encoder.")}, strict=True)
    enc.eval()
    up = torch.nn.Upsample(scale_factor=8, mode="bilinear")
    with torch.no_grad():
        e = enc(x)
        eu = up(e)
    out["mid_enc_logits_fold1"] = e.numpy()
    out["mid_enc_mask_fold1"] = eu[0].max(0)[1].byte().numpy()[None]

    # ---- 5-fold softmax ensemble on two small crops (extension; composition of reference modules)
    ens = O.synth_crops("D2", 2, 64, 64, seed=31, sigma=2.0)
    out["ens_u8"] = ens
    acc = 0
    for k in range(1, 6):
        mean, std = O.FOLD_MEAN_STD[k]
        xk = torch.from_numpy(np.stack([ref_preprocess(c, mean, std) for c in ens]))
        with torch.no_grad():
            acc = acc + torch.softmax(models[k](xk), dim=1)
    acc = acc / 5.0
    out["ens_prob"] = acc.numpy()
    out["ens_mask"] = acc.max(1)[1].byte().numpy()
    np.savez_compressed(os.path.join(HERE, "espnet_golden.npz"), **out)

    # ---- IoU: the reference's own iouEval on seeded label maps -----------------------------
    rng = np.random.default_rng(5)
    gt = rng.integers(0, 5, (3, 40, 50))
    pr = np.where(rng.random((3, 40, 50)) < 0.8, gt, rng.integers(0, 4, (3, 40, 50)))
    ev = iouEval(5)
    for i in range(3):
        ev.addBatch(pr[i], gt[i])
    overall, per_acc, per_iou, miou = ev.getMetricRight()
    np.savez(os.path.join(HERE, "iou_golden.npz"), gt=gt, pred=pr, hist=ev.hist, overall=overall,
             per_acc=per_acc, per_iou=per_iou, miou=miou)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print("  %-28s %8d B" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
