"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the drop-in modules
and therefore through the C ABI, against the golden fixtures generated from the real reference and
against the CPU oracle on the same seeded inputs.  Tolerances: fp32 logits <= 1e-3 max-abs (north_star);
integer / index work bit-exact."""
import numpy as np
import pytest
import torch

from glomeruli_segmentation_b200 import ESPNet, ESPNet_Encoder, ESPNetEnsemble, FOLD_MEAN_STD, _lib, iouEval, wsi
from oracle import espnet_oracle as O
from oracle import wsi_oracle as W

pytestmark = pytest.mark.gpu
LOGIT_TOL = 1e-3       # north_star: fp32 logits within 1e-3 max-abs
DEV = "cuda:0"


def _model(sd, classes=5, p=2, q=8):
    m = ESPNet(classes, p, q)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


def _maxabs(a, b):
    return (torch.as_tensor(a).float().cpu() - torch.as_tensor(b).float().cpu()).abs().max().item()


@pytest.mark.parametrize("fold", [1, 3])
def test_logits_and_masks_match_reference_golden(golden, fold_sd, fold):
    m = _model(fold_sd(fold))
    x = torch.from_numpy(golden["small_x_fold%d" % fold]).to(DEV)
    y = m(x)
    assert y.shape == (3, 5, 64, 96) and y.dtype == torch.float32
    assert _maxabs(y, golden["small_logits_fold%d" % fold]) <= LOGIT_TOL
    mask = y.max(1)[1].byte().cpu().numpy()
    assert (mask == golden["small_mask_fold%d" % fold]).mean() >= 0.9999


def test_u8_path_is_bit_identical_to_f32_path(golden, fold_sd):
    """P0 fused into the stem: the normalisation must be bit-exact, so both input formats give identical logits."""
    m = _model(fold_sd(1))
    mean, std = FOLD_MEAN_STD[1]
    y32 = m(torch.from_numpy(golden["small_x_fold1"]).to(DEV))
    u8 = torch.from_numpy(golden["small_u8"]).to(DEV)
    lg = torch.empty_like(y32)
    mask = m.segment(u8, mean, std, logits=lg)
    assert torch.equal(lg, y32)
    assert torch.equal(mask, y32.max(1)[1].byte())
    assert (mask.cpu().numpy() == golden["small_mask_fold1"]).mean() >= 0.9999
    host = m.segment_host(golden["small_u8"], mean, std)
    assert np.array_equal(host, mask.cpu().numpy())


def test_mid_crop_stages_against_oracle_and_golden(golden, fold_sd):
    sd = fold_sd(1)
    m = _model(sd)
    mean, std = FOLD_MEAN_STD[1]
    xn = O.normalise_bgr_u8(golden["mid_u8"], mean, std)
    y = m(torch.from_numpy(xn).to(DEV))
    taps = {}
    ref = O.espnet_forward(sd, torch.from_numpy(xn), taps)
    for stage, tap in [("b1", "b1"), ("b2", "b2"), ("b3", "b3"), ("up_l3", "up_l3"), ("up_l2", "up_l2")]:
        got = m.read_stage(stage).cpu().reshape(taps[tap].shape)
        assert _maxabs(got, taps[tap]) <= 2e-4, stage
    assert _maxabs(y, ref) <= LOGIT_TOL
    assert _maxabs(y, golden["mid_logits_fold1"]) <= LOGIT_TOL
    assert _maxabs(m.read_stage("b3").cpu().reshape(1, 256, 24, 32), golden["mid_tap:encoder.b3"].astype(np.float32)) <= 3e-2
    assert (y.max(1)[1].byte().cpu().numpy() == golden["mid_mask_fold1"]).mean() >= 0.9999


def test_encoder_only_and_upsampled_mask(golden, fold_sd):
    sd = O.encoder_state_dict(fold_sd(1))
    e = ESPNet_Encoder(5, 2, 8)
    e.load_state_dict(sd, strict=True)
    e = e.to(DEV).eval()
    mean, std = FOLD_MEAN_STD[1]
    xn = torch.from_numpy(O.normalise_bgr_u8(golden["mid_u8"], mean, std)).to(DEV)
    y = e(xn)
    assert y.shape == (1, 5, 24, 32)
    assert _maxabs(y, golden["mid_enc_logits_fold1"]) <= LOGIT_TOL
    mask = e.segment_upsampled(torch.from_numpy(golden["mid_u8"]).to(DEV), mean, std)
    assert (mask.cpu().numpy() == golden["mid_enc_mask_fold1"]).mean() >= 0.9995


def test_five_fold_softmax_ensemble(golden, fold_sd):
    models = [_model(fold_sd(k)) for k in range(1, 6)]
    ens = ESPNetEnsemble(models, [FOLD_MEAN_STD[k] for k in range(1, 6)])
    mask, prob = ens.segment(torch.from_numpy(golden["ens_u8"]).to(DEV), return_prob=True)
    assert _maxabs(prob, golden["ens_prob"]) <= 1e-4
    assert (mask.cpu().numpy() == golden["ens_mask"]).mean() >= 0.9995


@pytest.mark.parametrize("classes,p,q,B,H,W", [(5, 1, 1, 2, 8, 8), (5, 3, 2, 3, 16, 24), (5, 2, 3, 2, 72, 40),
                                               (20, 2, 3, 1, 40, 136), (5, 2, 8, 1, 264, 328)])
def test_random_weights_and_ragged_shapes_against_oracle(classes, p, q, B, H, W):
    sd = O.random_state_dict(classes, p, q, seed=100 + H)
    m = _model(sd, classes, p, q)
    x = torch.from_numpy(O.normalise_bgr_u8(O.synth_crops("D1", B, H, W, seed=H + W), *FOLD_MEAN_STD[2]))
    ref = O.espnet_forward(sd, x)
    y = m(x.to(DEV))
    assert y.shape == ref.shape
    assert _maxabs(y, ref) <= LOGIT_TOL * max(1.0, ref.abs().max().item() / 10.0)
    esd = O.encoder_state_dict(sd)
    e = ESPNet_Encoder(classes, p, q)
    e.load_state_dict(esd)
    e = e.to(DEV).eval()
    eref = O.espnet_encoder_forward(esd, x)
    assert _maxabs(e(x.to(DEV)), eref) <= LOGIT_TOL * max(1.0, eref.abs().max().item() / 10.0)


@pytest.mark.parametrize("impl", [0, 1, 2])
@pytest.mark.parametrize("B,H,W", [(1, 8, 8), (2, 72, 40), (1, 264, 328), (3, 256, 256)])
def test_all_fp32_branch_implementations_against_oracle(fold_sd, impl, B, H, W):
    """impl 0 = default (tcgen05, 3-term fp16 operand splits); CUDA-core kernels: 1 = per-thread global loads,
    2 = TMA-staged halo tiles.  All must hit the fp32 bar on ragged maps (level-3 maps down to 1x1 and widths that
    are not multiples of 4, 8 or 32)."""
    sd = fold_sd(1)
    m = _model(sd)
    if impl:
        m.set_option("fp32_impl", 0).set_option("branch_impl", impl)
    x = torch.from_numpy(O.normalise_bgr_u8(O.synth_crops("D2", B, H, W, seed=H, sigma=3.0), *FOLD_MEAN_STD[1]))
    ref = O.espnet_forward(sd, x)
    y = m(x.to(DEV))
    assert _maxabs(y, ref) <= LOGIT_TOL
    assert (y.max(1)[1].byte().cpu().numpy() == O.argmax_mask(ref)).mean() >= 0.9999


def test_full_size_crop_against_oracle_and_batch_invariance(fold_sd):
    sd = fold_sd(3)
    m = _model(sd)
    mean, std = FOLD_MEAN_STD[3]
    u8 = np.concatenate([O.synth_crops("D1", 1, 512, 512, seed=1), O.synth_crops("D2", 2, 512, 512, seed=2, sigma=8.0)])
    ref = O.espnet_forward(sd, torch.from_numpy(O.normalise_bgr_u8(u8[:2], mean, std)))
    d = torch.from_numpy(u8).to(DEV)
    lg = torch.empty((3, 5, 512, 512), device=DEV)
    mask = m.segment(d, mean, std, logits=lg)
    assert _maxabs(lg[:2], ref) <= LOGIT_TOL
    assert (mask[:2].cpu().numpy() == O.argmax_mask(ref)).mean() >= 0.9999
    # size-independent properties at the full size: a crop's result does not depend on its batch, and
    # two runs are bit-identical (no atomics / no order dependence in the forward)
    big = d.repeat(22, 1, 1, 1)[:64]
    mb = m.segment(big, mean, std)
    assert torch.equal(mb[0], mask[0]) and torch.equal(mb[3 * 21], mask[0]) and torch.equal(mb[62], mask[62 % 3])
    assert torch.equal(m.segment(big, mean, std), mb)


def test_error_behaviour(fold_sd):
    m = _model(fold_sd(1))
    with pytest.raises(RuntimeError, match="multiples of 8"):
        m(torch.zeros(1, 3, 20, 16, device=DEV))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 16, 16, device=DEV))
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.zeros(1, 3, 16, 16, device=DEV))
    m.eval()
    # editing weights + repack is honoured (no stale packed copy)
    x = torch.randn(1, 3, 16, 16, device=DEV) * 0.01
    y0 = m(x)
    sd = {k: v.clone() for k, v in fold_sd(1).items()}
    sd["classifier.weight"] = sd["classifier.weight"] * 2.0
    m.load_state_dict(sd)
    assert _maxabs(m(x), 2.0 * y0) <= 1e-4


# ------------------------------------------------------------------------------------------- WSI
def _synth_slide(h, w, seed):
    rng = np.random.default_rng(seed)
    coarse = rng.integers(60, 256, (h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
    up = np.kron(coarse, np.ones((16, 16, 1), np.float32))[:h, :w]
    return np.clip(up + rng.normal(0, 12, (h, w, 3)), 0, 255).astype(np.uint8)


def test_tiles_read_from_resident_slide_match_cut_tiles(fold_sd):
    m = _model(fold_sd(1))
    mean, std = FOLD_MEAN_STD[1]
    slide = _synth_slide(300, 420, 7)
    grid = wsi.tile_grid(420, 300, 96, 1.0, 1.0, 0.25, 1.0)      # last row / column overhang the slide
    org = grid.origins()
    cut = np.stack([W.read_tile(slide, int(x0), int(y0), grid.win_x, grid.win_y) for x0, y0 in org])
    a = m.segment(torch.from_numpy(cut).to(DEV), mean, std)
    b = m.segment_tiles(torch.from_numpy(slide).to(DEV), torch.from_numpy(org).to(DEV), grid.win_y, grid.win_x, mean, std)
    assert torch.equal(a, b)


@pytest.mark.parametrize("sw,sh,ws", [(1000, 760, 240), (500, 1300, 160), (333, 257, 80)])
def test_stitch_boxes_bit_exact(sw, sh, ws):
    rng = np.random.default_rng(sw)
    boxes, masks = [], []
    for _ in range(60):
        x0, y0 = int(rng.integers(-40, sw - 5)), int(rng.integers(-40, sh - 5))
        w, h = int(rng.integers(3, 220)), int(rng.integers(3, 220))
        boxes.append([float(x0), float(y0), float(x0 + w), float(y0 + h), 0.9])
        masks.append(rng.integers(0, 5, (h, w)).astype(np.uint8))
    level0, ds8 = W.stitch_slide(boxes, masks, sw, sh, ws)
    got = torch.zeros((sh, sw), dtype=torch.uint8, device=DEV)
    wsi.stitch_boxes(got, boxes, [torch.from_numpy(m).to(DEV) for m in masks], ws)
    assert np.array_equal(got.cpu().numpy(), level0)
    assert np.array_equal(wsi.downsample8(got, ws).cpu().numpy(), ds8)


@pytest.mark.parametrize("sw,sh,ov", [(1000, 760, 0.1), (640, 900, 0.5), (515, 389, 0.3)])
def test_stitch_grid_bit_exact_and_band_sharding(sw, sh, ov):
    ws = 80
    grid = wsi.tile_grid(sw, sh, 96, 1.0, 1.0, ov, 1.0)
    rng = np.random.default_rng(sh)
    tiles = rng.integers(0, 5, (grid.count, grid.win_y, grid.win_x)).astype(np.uint8)
    org = grid.origins()
    boxes = [[float(x), float(y), float(x + grid.win_x), float(y + grid.win_y), 1.0] for x, y in org]
    level0, ds8 = W.stitch_slide(boxes, list(tiles), sw, sh, ws)
    d_tiles = torch.from_numpy(tiles).to(DEV)
    full = torch.zeros((sh, sw), dtype=torch.uint8, device=DEV)
    wsi.stitch_grid(full, d_tiles, grid, 0, grid.n_y, ws)
    assert np.array_equal(full.cpu().numpy(), level0)
    assert np.array_equal(wsi.downsample8(full, ws).cpu().numpy(), ds8)
    # 3 "ranks": per-band stitch + element-wise max == the single pass (what the NCCL max-reduce does)
    acc = torch.zeros_like(full)
    for r in range(3):
        r0, rows = wsi.shard_rows(grid.n_y, r, 3)
        part = torch.zeros_like(full)
        wsi.stitch_grid(part, d_tiles[r0 * grid.n_x:(r0 + rows) * grid.n_x], grid, r0, rows, ws)
        acc = torch.maximum(acc, part)
    assert torch.equal(acc, full)


def test_small_slide_end_to_end_against_oracle_composition(fold_sd):
    """Config 4 in miniature: T1 tiles -> forward -> arg-max -> T3 -> T4, oracle = same composition on CPU."""
    sd = fold_sd(1)
    m = _model(sd)
    mean, std = FOLD_MEAN_STD[1]
    sh, sw, ws = 264, 392, 80
    slide = _synth_slide(sh, sw, 3)
    o, nx, ny, wx, wy, sx, sy = W.tile_grid(sw, sh, 128, 1.0, 1.0, 0.25, 1.0)
    cut = np.stack([W.read_tile(slide, int(x0), int(y0), wx, wy) for x0, y0 in o])
    ref_masks = O.argmax_mask(O.espnet_forward(sd, torch.from_numpy(O.normalise_bgr_u8(cut, mean, std))))
    boxes = [[float(x), float(y), float(x + wx), float(y + wy), 1.0] for x, y in o]
    ref0, ref8 = W.stitch_slide(boxes, list(ref_masks), sw, sh, ws)
    lvl0, ds8, n = wsi.segment_slide(m, torch.from_numpy(slide).to(DEV), mean, std, std_size=128, overlap=0.25, ws=ws, batch=5)
    assert n == nx * ny
    assert (lvl0.cpu().numpy() == ref0).mean() >= 0.9999
    assert (ds8.cpu().numpy() == ref8).mean() >= 0.9995


def test_gpu_confusion_histogram_matches_reference_ioueval():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "iou_golden.npz"))
    ev = iouEval(5)
    for i in range(3):
        ev.addBatch(torch.from_numpy(z["pred"][i].astype(np.uint8)).to(DEV), torch.from_numpy(z["gt"][i].astype(np.uint8)).to(DEV))
    assert np.array_equal(ev.hist, z["hist"])
    overall, per_acc, per_iou, miou = ev.getMetricRight()
    assert np.allclose(per_iou, z["per_iou"]) and np.isclose(miou, z["miou"])
    big_p = torch.randint(0, 5, (3_000_000,), dtype=torch.uint8, device=DEV)
    big_g = torch.randint(0, 5, (3_000_000,), dtype=torch.uint8, device=DEV)
    ev2 = iouEval(5)
    assert np.array_equal(ev2.addBatch(big_p, big_g), W.fast_hist(big_g.cpu().numpy(), big_p.cpu().numpy(), 5))


@pytest.mark.parametrize("net", ["full", "encoder"])
def test_unaligned_caller_buffers(fold_sd, net):
    """Input crops and output masks that are byte-offset views (no 4 / 8 / 16 B alignment): the vectorised stores fall back
    to scalar ones and the stem's byte loads do not care -- same masks as with aligned buffers."""
    from glomeruli_segmentation_b200 import ESPNet_Encoder
    sd = fold_sd(1)
    mean, std = FOLD_MEAN_STD[1]
    if net == "full":
        m = ESPNet(5, 2, 8); m.load_state_dict(sd, strict=True)
    else:
        m = ESPNet_Encoder(5, 2, 8); m.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    m = m.to(DEV).eval()
    B, H, W = 2, 72, 104
    u8 = torch.from_numpy(O.synth_crops("D2", B, H, W, seed=77, sigma=3.0)).to(DEV)
    lg_shape = (B, 5, H, W) if net == "full" else (B, 5, H // 8, W // 8)      # the encoder returns 1/8-resolution logits
    ref_lg = torch.empty(lg_shape, device=DEV)
    ref = m.segment(u8, mean, std, logits=ref_lg).clone()
    for off in (1, 2, 3, 5):
        raw_in = torch.empty(u8.numel() + 8, dtype=torch.uint8, device=DEV)
        vin = raw_in[off:off + u8.numel()].view(B, H, W, 3)
        vin.copy_(u8)
        raw_out = torch.zeros(B * H * W + 8, dtype=torch.uint8, device=DEV)
        vout = raw_out[off:off + B * H * W].view(B, H, W)
        assert vin.data_ptr() % 4 != 0 or off % 4 == 0
        raw_lg = torch.zeros(ref_lg.numel() + 8, device=DEV)
        vlg = raw_lg[off:off + ref_lg.numel()].view(lg_shape)          # 4-byte aligned only for odd `off`
        m.segment(vin, mean, std, out=vout, logits=vlg)
        assert torch.equal(vout, ref), off
        # the scalar fall-back kernels sum in a different order than the vectorised ones: fp32 rounding differences only
        assert (vlg - ref_lg).abs().max().item() <= 1e-4, off
        assert float(raw_lg[:off].abs().sum()) == 0 and float(raw_lg[off + ref_lg.numel():].abs().sum()) == 0
        assert int(raw_out[:off].sum()) == 0 and int(raw_out[off + B * H * W:].sum()) == 0    # nothing written outside the view


def _cpu_grid_stitch(sw, sh, grid, tiles, row0, rows, y_limit):
    out = np.zeros((sh, sw), np.uint8)
    for idx, (x0, y0) in enumerate(grid.origins(row0, rows)):
        m = tiles[idx]
        x1, y1 = min(x0 + grid.win_x, sw), min(min(y0 + grid.win_y, sh), y_limit)
        if x1 > x0 and y1 > y0:
            out[y0:y1, x0:x1] = np.maximum(out[y0:y1, x0:x1], m[: y1 - y0, : x1 - x0])
    return out


@pytest.mark.parametrize("sw,ov", [(515, 0.3), (640, 0.5), (1000, 0.25)])
def test_band_buffers_and_overlap_merge_equal_the_single_pass(sw, ov):
    """What a multi-GPU run does per rank (wsi.segment_slide): stitch ONLY the band's rows into a band-sized buffer, then the
    bands are placed / their overlap rows max-merged (espnet_max_merge_u8) -- emulated on one GPU for 3 'ranks'.  Widths and
    overlaps cover the byte path (width 515) and the aligned 32-bit-load path (width, window and stride all multiples of 4:
    640 at stride 48, 1000 at stride 72); the four-pixel path with unaligned tiles is test_stitch_grid_bit_exact's (1000, 86)."""
    sh, ws = 1389, 80
    grid = wsi.tile_grid(sw, sh, 96, 1.0, 1.0, ov, 1.0)
    rng = np.random.default_rng(5)
    tiles = rng.integers(0, 5, (grid.count, grid.win_y, grid.win_x)).astype(np.uint8)
    d_tiles = torch.from_numpy(tiles).to(DEV)
    full = torch.zeros((sh, sw), dtype=torch.uint8, device=DEV)
    wsi.stitch_grid(full, d_tiles, grid, 0, grid.n_y, ws)
    assert np.array_equal(full.cpu().numpy(), _cpu_grid_stitch(sw, sh, grid, tiles, 0, grid.n_y, wsi.stitch_y_limit(sw, sh, ws)))
    for balance in ("rows", "tiles"):
        out = torch.zeros_like(full)
        cov = 0
        for r in range(3):
            k0, k1, y0, y1 = wsi.band_tiles(grid, sh, r, 3, balance)
            band = torch.zeros((y1 - y0, sw), dtype=torch.uint8, device=DEV)
            wsi.stitch_grid(band, d_tiles[k0:k1], grid, 0, 0, ws, band_y0=y0, slide_h=sh, tiles=(k0, k1))
            # overwrite form into a dirty buffer: every pixel of the band's rows is written, nothing read
            dirty = torch.full_like(band, 9)
            wsi.stitch_grid(dirty, d_tiles[k0:k1], grid, 0, 0, ws, band_y0=y0, slide_h=sh, tiles=(k0, k1), overwrite=True)
            lim = wsi.stitch_y_limit(sw, sh, ws)
            assert torch.equal(dirty[:max(0, min(lim, y1) - y0)], band[:max(0, min(lim, y1) - y0)])
            split = min(max(cov, y0), y1)
            if split > y0:
                wsi.max_merge_(out[y0:split], band[:split - y0].contiguous())
            out[split:y1].copy_(band[split - y0:])
            cov = max(cov, y1)
        assert torch.equal(out, full), balance
    # unaligned / odd-length merge goes through the byte path
    a = torch.randint(0, 5, (1001,), dtype=torch.uint8, device=DEV)
    b = torch.randint(0, 5, (1001,), dtype=torch.uint8, device=DEV)
    exp = torch.maximum(a[1:], b[:-1])
    wsi.max_merge_(a[1:], b[:-1])
    assert torch.equal(a[1:], exp)


def test_stitch_grid_slide_taller_than_65535_rows():
    """a level-0 slide taller than 65535 px (common for 40x NDPI) used to be refused by the grid kernel."""
    sw, sh, ws = 96, 70001, 80
    grid = wsi.tile_grid(sw, sh, 64, 1.0, 1.0, 0.1, 1.0)
    rng = np.random.default_rng(6)
    coarse = rng.integers(0, 5, (grid.count, grid.win_y // 8, grid.win_x // 8)).astype(np.uint8)
    tiles = np.kron(coarse, np.ones((1, 8, 8), np.uint8))
    got = torch.zeros((sh, sw), dtype=torch.uint8, device=DEV)
    wsi.stitch_grid(got, torch.from_numpy(tiles).to(DEV), grid, 0, grid.n_y, ws)
    exp = _cpu_grid_stitch(sw, sh, grid, tiles, 0, grid.n_y, wsi.stitch_y_limit(sw, sh, ws))
    assert np.array_equal(got.cpu().numpy(), exp)
    assert exp[:wsi.stitch_y_limit(sw, sh, ws)].any() and not exp[wsi.stitch_y_limit(sw, sh, ws):].any()   # width quirk: only the first rows


def test_stitch_boxes_never_touches_bytes_past_an_odd_sized_mask():
    """SH*SW not a multiple of 4 -- the mask is a view whose end is followed by live data; those bytes survive."""
    sh, sw = 31, 33                                            # 1023 bytes; wider than tall, so the reference's ymax > width skip stays out of it
    buf = torch.full((sh * sw + 9,), 77, dtype=torch.uint8, device=DEV)
    view = buf[4:4 + sh * sw].view(sh, sw)
    view.zero_()
    rng = np.random.default_rng(2)
    boxes = [[20, 25, 40, 40, 1.0], [0, 28, 33, 31, 1.0], [30, 29, 33, 31, 1.0]]
    masks = [rng.integers(1, 5, (int(b[3] - b[1]), int(b[2] - b[0]))).astype(np.uint8) for b in boxes]
    wsi.stitch_boxes(view, boxes, [torch.from_numpy(m).to(DEV) for m in masks], ws=8)
    exp, _ = W.stitch_slide(boxes, masks, sw, sh, 8)
    assert np.array_equal(view.cpu().numpy(), exp) and exp[-1, -3:].all()
    assert bool((buf[:4] == 77).all()) and bool((buf[4 + sh * sw:] == 77).all())


def test_wrappers_reject_wrong_tensors():
    g = wsi.tile_grid(100, 100, 64, 1.0, 1.0, 0.1, 1.0)
    m = torch.zeros((100, 100), dtype=torch.uint8, device=DEV)
    with pytest.raises(RuntimeError):
        wsi.stitch_grid(m.t(), torch.zeros((g.count, 64, 64), dtype=torch.uint8, device=DEV), g, 0, g.n_y)      # not contiguous
    with pytest.raises(RuntimeError):
        wsi.stitch_grid(m, torch.zeros((g.count, 64, 64), dtype=torch.int32, device=DEV), g, 0, g.n_y)          # wrong dtype
    with pytest.raises(RuntimeError):
        wsi.stitch_grid(m, torch.zeros((g.count - 1, 64, 64), dtype=torch.uint8, device=DEV), g, 0, g.n_y)     # wrong tile count
    with pytest.raises(RuntimeError):
        wsi.stitch_boxes(m, [[0, 0, 10, 10, 1.0]], [torch.zeros((10, 10), dtype=torch.uint8)])                  # mask on the CPU


@pytest.mark.parametrize("net", ["full", "encoder"])
def test_captured_graph_replays_the_same_forward(fold_sd, net):
    """espnet_graph_capture / espnet_graph_launch: the batch-1 per-crop loop (VisualizeResults_iou.py:100-129) as one launch;
    masks and logits bit-equal to the plain forward, for several inputs through the same graph."""
    from glomeruli_segmentation_b200 import ESPNet_Encoder
    sd = fold_sd(1)
    mean, std = FOLD_MEAN_STD[1]
    if net == "full":
        m = ESPNet(5, 2, 8); m.load_state_dict(sd, strict=True)
    else:
        m = ESPNet_Encoder(5, 2, 8); m.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    m = m.to(DEV).eval()
    g = m.capture(1, 256, 256, mean, std, want_logits=True)
    l0 = _lib.lib().espnet_launch_count()
    for seed in (1, 2, 3):
        u8 = torch.from_numpy(O.synth_crops("D2", 1, 256, 256, seed=seed, sigma=3.0)).to(DEV)
        got = g.run(u8).clone()
        got_lg = g.logits.clone()
        ref_lg = torch.empty_like(got_lg)
        ref = m.segment(u8, mean, std, logits=ref_lg)
        assert torch.equal(got, ref) and torch.equal(got_lg, ref_lg)
    # 3 graph launches + 3 plain forwards: a graph replay counts as ONE launch
    per_forward = (_lib.lib().espnet_launch_count() - l0 - 3) // 3
    assert per_forward >= 20
    sd2 = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_state_dict(sd2)
    with pytest.raises(RuntimeError, match="capture again"):
        g.run()


@pytest.mark.parametrize("classes", [1, 2, 7, 20, 33])
@pytest.mark.parametrize("net", ["full", "encoder"])
def test_any_number_of_classes(classes, net):
    """Model.py:246,311 take any `classes`; 5 and 20 run the specialised tail kernels, the rest the generic run-time ones
    (kernels_tail_generic.cuh).  Logits against the oracle, masks = arg-max of the logits with the lowest-index tie rule."""
    from glomeruli_segmentation_b200 import ESPNet_Encoder
    sd = O.random_state_dict(classes, 2, 2, seed=classes)
    mean, std = FOLD_MEAN_STD[2]
    u8 = O.synth_crops("D2", 2, 64, 88, seed=classes, sigma=3.0)
    x = torch.from_numpy(O.normalise_bgr_u8(u8, mean, std))
    if net == "full":
        m = ESPNet(classes, 2, 2); m.load_state_dict(sd, strict=True)
        ref = O.espnet_forward(sd, x)
        lg = torch.empty((2, classes, 64, 88), device=DEV)
    else:
        esd = O.encoder_state_dict(sd)
        m = ESPNet_Encoder(classes, 2, 2); m.load_state_dict(esd, strict=True)
        ref = O.espnet_encoder_forward(esd, x)
        lg = torch.empty((2, classes, 8, 11), device=DEV)
    m = m.to(DEV).eval()
    mask = m.segment(torch.from_numpy(u8).to(DEV), mean, std, logits=lg)
    scale = max(1.0, ref.abs().max().item())
    assert (lg.cpu() - ref).abs().max().item() <= 1e-3 * scale
    assert (m(x.to(DEV)).cpu() - ref).abs().max().item() <= 1e-3 * scale
    if net == "full":
        assert torch.equal(mask, lg.max(1)[1].to(torch.uint8))
    else:
        up = O.upsample8_bilinear(lg.cpu())
        assert (mask.cpu().numpy() == O.argmax_mask(up)).mean() >= 0.999       # bilinear weights differ in the last ulp only at ties


@pytest.mark.parametrize("net", ["full", "encoder"])
def test_generic_tail_is_bit_identical_to_the_specialised_scalar_tail(fold_sd, net):
    from glomeruli_segmentation_b200 import ESPNet_Encoder
    sd = fold_sd(2)
    mean, std = FOLD_MEAN_STD[2]
    if net == "full":
        mk = lambda: ESPNet(5, 2, 8)
        state = sd
    else:
        mk = lambda: ESPNet_Encoder(5, 2, 8)
        state = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    a, b = mk(), mk()
    a.load_state_dict(state, strict=True); b.load_state_dict(state, strict=True)
    a = a.to(DEV).eval().set_option("dec_impl", 0)            # specialised one-pixel-per-thread kernels
    b = b.to(DEV).eval().set_option("tail_impl", 1)           # generic run-time-nc kernels
    u8 = torch.from_numpy(O.synth_crops("D2", 2, 136, 200, seed=4, sigma=3.0)).to(DEV)
    shape = (2, 5, 136, 200) if net == "full" else (2, 5, 17, 25)
    la, lb = torch.empty(shape, device=DEV), torch.empty(shape, device=DEV)
    ma = a.segment(u8, mean, std, logits=la)
    mb = b.segment(u8, mean, std, logits=lb)
    assert torch.equal(la, lb)
    if net == "full":
        assert torch.equal(ma, mb)
    else:       # the generic x8 up-sampler evaluates the same bilinear formula per pixel instead of per 4-pixel group
        assert (ma != mb).sum().item() <= 1e-4 * ma.numel()


@pytest.mark.parametrize("mode", ["fp32", "f16tc"])
@pytest.mark.parametrize("net", ["full", "encoder"])
def test_programmatic_dependent_launch_never_changes_a_result(fold_sd, net, mode):
    """"pdl" option: every kernel of the forward may become resident while the one before it drains (it waits before touching
    activations).  Off, automatic, every class on and two partial masks give bit-identical logits and masks, plain launches
    and graph replay, also when forwards follow each other back to back on one stream (the next stem behind the last kernel)."""
    from glomeruli_segmentation_b200 import ESPNet_Encoder
    sd = fold_sd(3)
    mean, std = FOLD_MEAN_STD[3]
    if net == "full":
        m = ESPNet(5, 2, 8); m.load_state_dict(sd, strict=True)
    else:
        m = ESPNet_Encoder(5, 2, 8); m.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    m = m.to(DEV).eval().set_mode(mode)
    u8 = torch.from_numpy(O.synth_crops("D2", 3, 264, 200, seed=9, sigma=3.0)).to(DEV)
    shape = (3, 5, 264, 200) if net == "full" else (3, 5, 33, 25)
    want_lg = torch.empty(shape, device=DEV)
    m.set_option("pdl", 0)
    want = m.segment(u8, mean, std, logits=want_lg).clone()
    for mask in (-1, 127, 126, 56, 7):
        m.set_option("pdl", mask)
        lg = [torch.empty(shape, device=DEV) for _ in range(4)]
        got = [m.segment(u8, mean, std, logits=lg[i]) for i in range(4)]          # four forwards enqueued back to back
        for i in range(4):
            assert torch.equal(got[i], want) and torch.equal(lg[i], want_lg), (mask, i)
        g = m.capture(3, 264, 200, mean, std, want_logits=True)
        for _ in range(3):
            g.run(u8)
        assert torch.equal(g.mask, want) and torch.equal(g.logits, want_lg), mask
    with pytest.raises(RuntimeError):
        m.set_option("pdl", 128)


def test_host_pipelines_match_the_resident_calls(fold_sd):
    """HostPipeline (pinned host crops in, host masks out, copies overlapped with the kernels of neighbouring batches) for a
    single model and for the 5-fold ensemble: every batch of a stream of different batches equals the plain segment() call."""
    models = [_model(fold_sd(k)) for k in range(1, 6)]
    ens = ESPNetEnsemble(models, [FOLD_MEAN_STD[k] for k in range(1, 6)])
    mean, std = FOLD_MEAN_STD[1]
    B, H, W = 3, 72, 104
    batches = [torch.from_numpy(O.synth_crops("D2", B, H, W, seed=40 + i, sigma=3.0)).pin_memory() for i in range(5)]
    for owner, plain, pipe in ((models[0], lambda d: models[0].segment(d, mean, std), models[0].host_pipeline(B, H, W, mean, std, depth=2)),
                               (ens, lambda d: ens.segment(d), ens.host_pipeline(B, H, W, depth=2))):
        outs = [torch.empty((B, H, W), dtype=torch.uint8).pin_memory() for _ in batches]
        for u8, o in zip(batches, outs):
            pipe.submit(u8, o)
        pipe.drain()
        for u8, o in zip(batches, outs):
            assert torch.equal(o, plain(u8.to(DEV)).cpu())
    with pytest.raises(RuntimeError):
        ens.segment(batches[0].to(DEV), out=torch.empty((B, H, W), dtype=torch.float32, device=DEV))
