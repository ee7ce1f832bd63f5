"""CPU oracle for the ESPNet inference hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (state_dict-driven, functional) of the reference network
`module/espnet/test/Model.py` of jinseikenai/glomeruli_segmentation.  It is the checker for
the CUDA path; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  The product (`glomeruli_segmentation_b200`) never does.

Where the arithmetic lives: the reference has no kernels of its own, every op is a call into
third-party PyTorch (reference pin: unversioned conda "PyTorch 1.1", `docker/gpu.dockerfile:38-41`;
this image: torch 2.11.0 CPU/oneDNN).  The restatement therefore calls the same
`torch.nn.functional` ops on CPU, in the same order, and is pinned against outputs of the real
reference `Model.py` + shipped `models/espnet_fold*.pth` run in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`): parity pinned by generated fixtures,
the reference itself holds no tests.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-3  # Model.py:21,44,69,142,331  (every BatchNorm2d uses eps=1e-03)

# Per-fold BGR mean / std published by the reference (README.md:243-249).
FOLD_MEAN_STD = {
    1: ((204.60071, 170.19359, 199.57469), (20.61257, 42.92207, 28.401505)),
    2: ((202.38148, 167.13171, 198.10599), (20.704079, 42.958416, 28.366297)),
    3: ((203.12099, 167.813, 198.50894), (21.038654, 43.769535, 29.034416)),
    4: ((203.66399, 167.94217, 198.58081), (20.96783, 43.556736, 28.838718)),
    5: ((204.49896, 169.03307, 199.22058), (20.547842, 42.86628, 27.966227)),
}


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def _bn(sd: Dict[str, Tensor], key: str, x: Tensor) -> Tensor:
    """Eval-mode BatchNorm2d(eps=1e-3): (x-mean)*gamma/sqrt(var+eps)+beta."""
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                        sd[key + ".weight"], sd[key + ".bias"], False, 0.0, BN_EPS)


def _prelu(sd: Dict[str, Tensor], key: str, x: Tensor) -> Tensor:
    return F.prelu(x, sd[key + ".weight"])


def _conv(sd: Dict[str, Tensor], key: str, x: Tensor, stride: int = 1, dilation: int = 1) -> Tensor:
    """Bias-free k x k conv with 'same'-style padding ((k-1)/2)*d  (Model.py:19-20,95-96,119-120)."""
    w = sd[key + ".conv.weight"]
    pad = ((w.shape[-1] - 1) // 2) * dilation
    return F.conv2d(x, w, None, stride, pad, dilation)


def cbr(sd, key, x, stride=1):
    """CBR: conv -> BN -> PReLU  (Model.py:6-32)."""
    return _prelu(sd, key + ".act", _bn(sd, key + ".bn", _conv(sd, key, x, stride)))


def br(sd, key, x):
    """BR: BN -> PReLU  (Model.py:35-54)."""
    return _prelu(sd, key + ".act", _bn(sd, key + ".bn", x))


def _branches(sd, key, o1):
    """Five dilated 3x3 branches + hierarchical feature fusion + concat
    (Model.py:146-157 and :198-208).  Channel order: d1, add1, add2, add3, add4."""
    d1 = _conv(sd, key + ".d1", o1, 1, 1)
    s = _conv(sd, key + ".d2", o1, 1, 2)
    parts = [d1, s]
    for d in (4, 8, 16):
        s = s + _conv(sd, key + ".d%d" % d, o1, 1, d)
        parts.append(s)
    return torch.cat(parts, 1)


def down_sampler(sd, key, x, taps=None):
    """DownSamplerB (Model.py:130-160): 3x3 stride-2 reduce, branches, BN, PReLU (no residual)."""
    o1 = _conv(sd, key + ".c1", x, 2)
    if taps is not None:
        taps[key + ".c1"] = o1
    y = _branches(sd, key, o1)
    return _prelu(sd, key + ".act", _bn(sd, key + ".bn", y))


def esp_block(sd, key, x, taps=None):
    """DilatedParllelResidualBlockB (Model.py:162-214): 1x1 reduce, branches, residual add
    *before* BN (:211-212), then BR."""
    o1 = _conv(sd, key + ".c1", x, 1)
    if taps is not None:
        taps[key + ".c1"] = o1
    y = x + _branches(sd, key, o1)
    return br(sd, key + ".bn", y)


def avg_pool(x):
    """InputProjectionA's AvgPool2d(3, stride=2, padding=1), count_include_pad (Model.py:230)."""
    return F.avg_pool2d(x, 3, 2, 1)


def conv_transpose(w: Tensor, x: Tensor) -> Tensor:
    """ConvTranspose2d(k=2, s=2, p=0, bias=False)  (Model.py:334,337,339)."""
    return F.conv_transpose2d(x, w, None, 2, 0)


# ----------------------------------------------------------------------------------------------
# networks
# ----------------------------------------------------------------------------------------------
def count_blocks(sd: Dict[str, Tensor], prefix: str):
    p = 0
    while (prefix + "level2.%d.c1.conv.weight" % p) in sd:
        p += 1
    q = 0
    while (prefix + "level3.%d.c1.conv.weight" % q) in sd:
        q += 1
    return p, q


def encoder_trunk(sd, x, prefix="", taps: Optional[dict] = None):
    """ESPNet_Encoder.forward up to output2_cat (Model.py:278-300) == ESPNet.forward :346-368."""
    e = prefix
    p, q = count_blocks(sd, e)
    t = taps if taps is not None else {}
    out0 = cbr(sd, e + "level1", x, 2)                       # :278 / :346
    inp1 = avg_pool(x)                                       # :279 / :347
    inp2 = avg_pool(inp1)                                    # :280 / :348
    out0_cat = br(sd, e + "b1", torch.cat([out0, inp1], 1))  # :282 / :350
    out1_0 = down_sampler(sd, e + "level2_0", out0_cat, taps)      # :283 / :351
    out1 = out1_0
    t["level1"], t["b1"], t["level2_0"] = out0, out0_cat, out1_0
    for i in range(p):                                       # :285-289
        out1 = esp_block(sd, e + "level2.%d" % i, out1, taps)
        t["level2.%d" % i] = out1
    out1_cat = br(sd, e + "b2", torch.cat([out1, out1_0, inp2], 1))   # :291 / :359
    out2_0 = down_sampler(sd, e + "level3_0", out1_cat, taps)         # :293 / :361
    out2 = out2_0
    t["b2"], t["level3_0"] = out1_cat, out2_0
    for i in range(q):                                       # :294-298
        out2 = esp_block(sd, e + "level3.%d" % i, out2, taps)
        t["level3.%d" % i] = out2
    out2_cat = br(sd, e + "b3", torch.cat([out2_0, out2], 1))         # :300 / :368
    t["b3"] = out2_cat
    return out0_cat, out1_cat, out2_cat


@torch.no_grad()
def espnet_encoder_forward(sd, x, prefix="", taps: Optional[dict] = None) -> Tensor:
    """ESPNet_Encoder.forward (Model.py:273-304): [B,3,H,W] -> [B,classes,H/8,W/8]."""
    _, _, out2_cat = encoder_trunk(sd, x, prefix, taps)
    y = F.conv2d(out2_cat, sd[prefix + "classifier.conv.weight"])     # :302
    if taps is not None:
        taps["encoder.classifier"] = y
    return y


@torch.no_grad()
def espnet_forward(sd, x, taps: Optional[dict] = None) -> Tensor:
    """ESPNet.forward (Model.py:341-378): [B,3,H,W] -> logits [B,classes,H,W]; H,W % 8 == 0."""
    t = taps if taps is not None else {}
    out0_cat, out1_cat, out2_cat = encoder_trunk(sd, x, "encoder.", taps)
    enc_cls = F.conv2d(out2_cat, sd["encoder.classifier.conv.weight"])
    out2_c = conv_transpose(sd["up_l3.0.weight"], _bn(sd, "br", enc_cls))       # :370
    out1_c = F.conv2d(out1_cat, sd["level3_C.conv.weight"])                       # :372
    z = br(sd, "combine_l2_l3.0", torch.cat([out1_c, out2_c], 1))                 # :373
    z = cbr(sd, "combine_l2_l3.1", z)
    comb = br(sd, "up_l2.1", conv_transpose(sd["up_l2.0.weight"], z))             # :373
    feat = cbr(sd, "conv", torch.cat([comb, out0_cat], 1))                        # :375
    logits = conv_transpose(sd["classifier.weight"], feat)                        # :377
    t["encoder.classifier"], t["up_l3"], t["level3_C"] = enc_cls, out2_c, out1_c
    t["combine_l2_l3"], t["up_l2"], t["conv"], t["classifier"] = z, comb, feat, logits
    return logits


# ----------------------------------------------------------------------------------------------
# pre / post processing around the forward
# ----------------------------------------------------------------------------------------------
def normalise_bgr_u8(img_u8: np.ndarray, mean, std) -> np.ndarray:
    """P0 -- VisualizeResults_iou.py:107-119 for a crop that already has the network size
    (cv2.resize at :114 is then the identity): three separate fp32 roundings
    ((u8 - mean_c) / std_c) / 255, c in BGR order, HWC -> CHW.  Accepts [H,W,3] or [B,H,W,3]."""
    a = img_u8.astype(np.float32)
    m = np.asarray(mean, dtype=np.float32)
    s = np.asarray(std, dtype=np.float32)
    a = a - m
    a = a / s
    a = a / np.float32(255)
    return np.ascontiguousarray(np.moveaxis(a, -1, -3))


def argmax_mask(logits: Tensor) -> np.ndarray:
    """A10 -- VisualizeResults_iou.py:128: per-pixel max over channels, ties -> lowest index, u8."""
    return logits.max(1)[1].to(torch.uint8).cpu().numpy()


def upsample8_bilinear(x: Tensor) -> Tensor:
    """ESPNet-C path: nn.Upsample(scale_factor=8, mode='bilinear') (VisualizeResults_iou.py:258-261,
    125-126); align_corners default False."""
    return F.interpolate(x, scale_factor=8, mode="bilinear", align_corners=False)


@torch.no_grad()
def ensemble_mask(sds, crops_u8: np.ndarray, folds) -> np.ndarray:
    """Extension: per fold k, x_k = P0(u8; mean_k, std_k), p_k = softmax(ESPNet_k(x_k));
    mask = argmax(mean_k p_k), ties -> lowest index.  Not present in the reference; a composition
    of reference modules."""
    acc = None
    for sd, k in zip(sds, folds):
        mean, std = FOLD_MEAN_STD[k]
        x = torch.from_numpy(normalise_bgr_u8(crops_u8, mean, std))
        pk = torch.softmax(espnet_forward(sd, x), dim=1)
        acc = pk if acc is None else acc + pk
    acc = acc / float(len(folds))
    return argmax_mask(acc), acc


# ----------------------------------------------------------------------------------------------
# synthetic inputs: D1 iid uniform, D2 blurred stain noise, D3 smooth
# ----------------------------------------------------------------------------------------------
def synth_crops(kind: str, n: int, h: int, w: int, seed: int = 0, sigma: float = 4.0) -> np.ndarray:
    """Seeded synthetic BGR u8 crops [n,h,w,3]."""
    out = np.empty((n, h, w, 3), np.uint8)
    for i in range(n):
        rng = np.random.default_rng(seed + i)
        if kind == "D1":
            out[i] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == "D2":
            from scipy.ndimage import gaussian_filter
            base = np.array([150.0, 120.0, 150.0])
            img = np.empty((h, w, 3))
            for c in range(3):
                nz = gaussian_filter(rng.standard_normal((h, w)), sigma, mode="wrap")
                img[..., c] = base[c] + 90.0 * nz / (nz.std() + 1e-12)
            out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        elif kind == "D3":
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([215 + 8 * np.sin(xx / 37.0 + i), 185 + 10 * np.cos(yy / 29.0), 210 + 6 * np.sin((xx + yy) / 53.0)], -1)
            out[i] = np.clip(np.rint(img + rng.normal(0, 1.0, (h, w, 3))), 0, 255).astype(np.uint8)
        else:
            raise ValueError(kind)
    return out


def random_state_dict(classes=5, p=2, q=8, seed=0, encoder_only=False) -> Dict[str, Tensor]:
    """Random-init weights of the ESPNet(classes,p,q) architecture with the reference's key set and
    checkpoint-like statistics (negative PReLU slopes, tiny running_var, negative gamma) --
    used where the shipped checkpoints are not available (bench on a box without fixtures)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(key, co, ci, k):
        sd[key + ".conv.weight"] = torch.randn(co, ci, k, k, generator=g) * math.sqrt(1.0 / (ci * k * k))

    def bn(key, c):
        gamma = (torch.randn(c, generator=g) * 0.25 + 0.7) * torch.where(torch.rand(c, generator=g) < 0.15, -1.0, 1.0)
        sd[key + ".bias"] = torch.randn(c, generator=g) * 0.2
        sd[key + ".running_mean"] = torch.randn(c, generator=g) * 0.2
        var = torch.rand(c, generator=g) * 1.0 + 0.5
        var[::17] = 5.6e-45          # dead channels of the shipped checkpoints: denormal variance ...
        gamma[::17] = gamma[::17] * 0.02   # ... whose 1/sqrt(eps) = 31x gain is tamed by a tiny gamma
        sd[key + ".weight"] = gamma
        sd[key + ".running_var"] = var
        sd[key + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)

    def act(key, c):
        sd[key + ".weight"] = torch.rand(c, generator=g) * 1.2 - 0.5

    def down(key, ci, co):
        n = co // 5
        n1 = co - 4 * n
        conv(key + ".c1", n, ci, 3)
        conv(key + ".d1", n1, n, 3)
        for d in (2, 4, 8, 16):
            conv(key + ".d%d" % d, n, n, 3)
        bn(key + ".bn", co)
        act(key + ".act", co)

    def esp(key, c):
        n = c // 5
        n1 = c - 4 * n
        conv(key + ".c1", n, c, 1)
        conv(key + ".d1", n1, n, 3)
        for d in (2, 4, 8, 16):
            conv(key + ".d%d" % d, n, n, 3)
        bn(key + ".bn.bn", c)
        act(key + ".bn.act", c)

    e = "" if encoder_only else "encoder."
    conv(e + "level1", 16, 3, 3); bn(e + "level1.bn", 16); act(e + "level1.act", 16)
    bn(e + "b1.bn", 19); act(e + "b1.act", 19)
    down(e + "level2_0", 19, 64)
    for i in range(p):
        esp(e + "level2.%d" % i, 64)
    bn(e + "b2.bn", 131); act(e + "b2.act", 131)
    down(e + "level3_0", 131, 128)
    for i in range(q):
        esp(e + "level3.%d" % i, 128)
    bn(e + "b3.bn", 256); act(e + "b3.act", 256)
    conv(e + "classifier", classes, 256, 1)
    if not encoder_only:
        conv("level3_C", classes, 131, 1)
        bn("br", classes)
        conv("conv", classes, 19 + classes, 3); bn("conv.bn", classes); act("conv.act", classes)
        sd["up_l3.0.weight"] = torch.randn(classes, classes, 2, 2, generator=g) * 0.4
        bn("combine_l2_l3.0.bn", 2 * classes); act("combine_l2_l3.0.act", 2 * classes)
        conv("combine_l2_l3.1", classes, 2 * classes, 3)
        bn("combine_l2_l3.1.bn", classes); act("combine_l2_l3.1.act", classes)
        sd["up_l2.0.weight"] = torch.randn(classes, classes, 2, 2, generator=g) * 0.4
        bn("up_l2.1.bn", classes); act("up_l2.1.act", classes)
        sd["classifier.weight"] = torch.randn(classes, classes, 2, 2, generator=g) * 0.4
    return sd


def encoder_state_dict(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """The `encoder.`-stripped subset that ESPNet_Encoder loads strictly."""
    return {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
