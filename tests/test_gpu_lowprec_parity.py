"""Reduced-precision path (ESPNET_MODE_F16TC) held to north_star's bar against the ORACLE (not against the repo's own fp32
mode): arg-max agreement >= 0.999 and set-accumulated IoU (iouEval.getMetricRight, IOUEval.py:63-69) per class with support,
on the distributions SURVEY.md 7.1 names as the hard ones (D1 iid-uniform, D2 blurred noise sigma 2 and 16), for ALL FIVE
shipped folds at 512 x 512; the 5-fold softmax ensemble of BASELINE configs[2] in that mode at 512 x 512; directed arg-max
tie tests (VisualizeResults_iou.py:128: ties -> lowest class index)."""
import numpy as np
import pytest
import torch

from glomeruli_segmentation_b200 import ESPNet, ESPNet_Encoder, ESPNetEnsemble, FOLD_MEAN_STD, iouEval
from oracle import espnet_oracle as O
from oracle import wsi_oracle as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
AGREE = 0.999
N_CROPS = 8


def _model(sd, mode):
    m = ESPNet(5, 2, 8)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval().set_mode(mode)


_ORACLE_CACHE = {}


def _oracle_masks(fold_sd, fold, dist, sigma):
    """Oracle (CPU restatement pinned to the reference by tests/golden) masks of the 8 seeded 512^2 crops."""
    key = (fold, dist, sigma)
    if key not in _ORACLE_CACHE:
        mean, std = FOLD_MEAN_STD[fold]
        u8 = O.synth_crops(dist, N_CROPS, 512, 512, seed=1000 * fold + int(sigma), sigma=sigma or 4.0)
        ref = torch.cat([O.espnet_forward(fold_sd(fold), torch.from_numpy(O.normalise_bgr_u8(u8[i:i + 4], mean, std)))
                         for i in range(0, N_CROPS, 4)])
        _ORACLE_CACHE[key] = (u8, O.argmax_mask(ref))
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("dist,sigma", [("D1", 0), ("D2", 2.0), ("D2", 16.0)])
@pytest.mark.parametrize("fold", [1, 2, 3, 4, 5])
def test_f16tc_masks_vs_oracle_all_folds_hard_distributions(fold_sd, fold, dist, sigma):
    u8, ref_mask = _oracle_masks(fold_sd, fold, dist, sigma)
    mean, std = FOLD_MEAN_STD[fold]
    m = _model(fold_sd(fold), "f16tc")
    mask = m.segment(torch.from_numpy(u8).to(DEV), mean, std)
    agree = float((mask.cpu().numpy() == ref_mask).mean())
    assert agree >= AGREE, (fold, dist, sigma, agree)
    # IoU through the drop-in iouEval (GPU histogram) == the oracle's histogram, then the reference's getMetricRight formula
    ev = iouEval(5)
    ev.addBatch(mask, torch.from_numpy(ref_mask).to(DEV))
    hist = np.asarray(ev.hist, np.float64)
    assert np.array_equal(hist, W.fast_hist(ref_mask, mask.cpu().numpy(), 5))
    _, _, per_iou, _ = ev.getMetricRight()
    n_pix = hist.sum()
    for c in range(5):
        union = hist[c, :].sum() + hist[:, c].sum() - hist[c, c]
        if union == 0:
            continue
        # north_star's 0.999 is a PIXEL budget: a class's IoU loss, (1 - IoU_c) * union_c pixels, must fit into the 0.1 % of
        # pixels allowed to differ -- which is the only IoU statement 0.999 agreement supports for a class with 1 % support --
        # and classes that carry >= 10 % of the pixels must hold IoU >= 0.995 outright (measured 0.9988-0.9998).
        assert (1.0 - per_iou[c]) * union <= (1.0 - AGREE) * n_pix, (fold, dist, sigma, c, per_iou[c], union / n_pix)
        if union >= 0.10 * n_pix:
            assert per_iou[c] >= 0.995, (fold, dist, sigma, c, per_iou[c])


@pytest.mark.parametrize("mode", ["f16tc", "fp32"])
def test_ensemble_512_vs_oracle(fold_sd, mode):
    """BASELINE configs[2] in miniature: 5 folds, per-fold mean/std, softmax mean, arg-max -- at 512 x 512, in the
    reduced-precision mode the config is quoted on (and in the fp32-equivalent mode)."""
    folds = [1, 2, 3, 4, 5]
    sds = [fold_sd(k) for k in folds]
    u8 = O.synth_crops("D2", 4, 512, 512, seed=77, sigma=6.0)
    ref_mask, ref_prob = O.ensemble_mask(sds, u8, folds)
    ens = ESPNetEnsemble([_model(sd, mode) for sd in sds], [FOLD_MEAN_STD[k] for k in folds])
    mask, prob = ens.segment(torch.from_numpy(u8).to(DEV), return_prob=True)
    agree = float((mask.cpu().numpy() == ref_mask).mean())
    assert agree >= (AGREE if mode == "f16tc" else 0.9999), agree
    assert (prob.cpu() - ref_prob).abs().max().item() <= (2e-2 if mode == "f16tc" else 2e-4)


def _tie_sd(net):
    """Random weights in which two pairs of classes (0,1) and (3,4) get IDENTICAL last-layer weights: their logits are
    bit-equal at every pixel, so the winner must be the lower index (VisualizeResults_iou.py:128, max(0)[1])."""
    sd = O.random_state_dict(5, 2, 3, seed=9)
    if net == "full":
        w = sd["classifier.weight"].clone()            # ConvTranspose2d weight [Cin, Cout, 2, 2]
        w[:, 1] = w[:, 0]
        w[:, 4] = w[:, 3]
        sd["classifier.weight"] = w
        return sd
    sd = O.encoder_state_dict(sd)
    w = sd["classifier.conv.weight"].clone()           # Conv2d weight [Cout, Cin, 1, 1]
    w[1] = w[0]
    w[4] = w[3]
    sd["classifier.conv.weight"] = w
    return sd


@pytest.mark.parametrize("mode", ["fp32", "f16tc"])
@pytest.mark.parametrize("net", ["full", "encoder"])
def test_argmax_ties_go_to_the_lowest_class(net, mode):
    sd = _tie_sd(net)
    m = (ESPNet(5, 2, 3) if net == "full" else ESPNet_Encoder(5, 2, 3))
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval().set_mode(mode)
    mean, std = FOLD_MEAN_STD[1]
    u8 = torch.from_numpy(O.synth_crops("D2", 2, 96, 136, seed=3, sigma=3.0)).to(DEV)
    lg_shape = (2, 5, 96, 136) if net == "full" else (2, 5, 12, 17)
    lg = torch.empty(lg_shape, device=DEV)
    mask = m.segment(u8, mean, std, logits=lg)
    assert torch.equal(lg[:, 0], lg[:, 1]) and torch.equal(lg[:, 3], lg[:, 4])      # the ties are real
    vals = set(np.unique(mask.cpu().numpy()).tolist())
    assert vals <= {0, 2, 3} and (vals & {0, 3}), vals                              # never 1 or 4, and a tied pair does win somewhere
    if net == "full":
        assert torch.equal(mask, lg.max(1)[1].to(torch.uint8))                     # torch's own tie rule, same logits
