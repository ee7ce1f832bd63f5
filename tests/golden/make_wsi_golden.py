#!/usr/bin/env python
"""Generate tests/golden/wsi_golden.npz and frontend_golden.npz by RUNNING THE REAL REFERENCE SCRIPTS
(unmodified, imported from /root/reference) in the build container:

  T1   module/faster-rcnn/detect_glomus_test.py   GlomusDetector.scan_region / calc_window_size  (:236-304)
  T2-4 module/espnet/test/eval_wsi_segmentation.py Generate_Segmentation_Gt.generate_pred_wsi, overlay,
       generate_whole_img  (:215-316, :359-394)   + module/common/annotation_handler.py check_overlap (:74-105)
  f2   module/faster-rcnn/make_seg_data.py         Generate_Segmentation_Gt.output_org_files (:347-361)
  P0/f2/A10/A11/f4  module/espnet/test/VisualizeResults_iou.py evaluateModel (:84-156) with the real Model.py,
       the shipped checkpoints and the stock cv2 of this image.

The scripts import tensorflow / openslide / labelme / matplotlib at module level; none of the code under test
uses tensorflow or matplotlib, so those are empty stub modules.  What IS stubbed with behaviour:
  * openslide.open_slide -> an in-memory slide (numpy RGB level 0; read_region pads out-of-bounds pixels with
    transparent black like OpenSlide documents) that records every read_region call;
  * labelme.utils.img_b64_to_arr / img_arr_to_b64 -> PIL PNG <-> numpy, labelme's published two-liners;
  * PIL.ImageFont.truetype falls back to the default font (DejaVuSans.ttf is not in this image; only used for drawing).
All index arithmetic, numpy slicing, cv2 calls and the network forward are the reference's own code.

Run:  python tests/golden/make_wsi_golden.py      (needs /root/reference; the fixtures are committed)
"""
import base64
import contextlib
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import wsi_cases as WC  # noqa: E402


# ------------------------------------------------------------------------------------------- stubs
class FakeSlide:
    """Stand-in for openslide.OpenSlide over an in-memory level-0 RGB array (or just dimensions)."""

    def __init__(self, width, height, rgb=None, mpp=(0.2277, 0.2277), power=40, level_downsamples=(1.0,)):
        self.dimensions = (width, height)
        self.rgb = rgb
        self.level_downsamples = tuple(level_downsamples)
        self.properties = {"openslide.mpp-x": str(mpp[0]), "openslide.mpp-y": str(mpp[1]),
                           "openslide.objective-power": str(power)}
        self.calls = []

    def read_region(self, location, level, size):
        from PIL import Image
        x0, y0 = int(location[0]), int(location[1])
        w, h = int(size[0]), int(size[1])
        self.calls.append((x0, y0, int(level), w, h))
        out = np.zeros((h, w, 4), np.uint8)
        if self.rgb is not None and level == 0:
            sh, sw = self.rgb.shape[:2]
            cx0, cy0, cx1, cy1 = max(x0, 0), max(y0, 0), min(x0 + w, sw), min(y0 + h, sh)
            if cx1 > cx0 and cy1 > cy0:
                out[cy0 - y0:cy1 - y0, cx0 - x0:cx1 - x0, :3] = self.rgb[cy0:cy1, cx0:cx1]
                out[cy0 - y0:cy1 - y0, cx0 - x0:cx1 - x0, 3] = 255
        return Image.fromarray(out, "RGBA")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


SLIDES = {}          # path -> FakeSlide, consulted by the stub open_slide


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("tensorflow")
    mod("openslide", open_slide=lambda p: SLIDES[p], PROPERTY_NAME_MPP_X="openslide.mpp-x",
        PROPERTY_NAME_MPP_Y="openslide.mpp-y", PROPERTY_NAME_OBJECTIVE_POWER="openslide.objective-power")

    def img_b64_to_arr(img_b64):                     # labelme/utils/image.py
        import PIL.Image
        f = io.BytesIO()
        f.write(base64.b64decode(img_b64))
        return np.array(PIL.Image.open(f))

    def img_arr_to_b64(img_arr):
        import PIL.Image
        f = io.BytesIO()
        PIL.Image.fromarray(img_arr).save(f, format="PNG")
        return base64.encodebytes(f.getvalue()) if hasattr(base64, "encodebytes") else base64.encodestring(f.getvalue())

    utils = mod("labelme.utils", img_b64_to_arr=img_b64_to_arr, img_arr_to_b64=img_arr_to_b64)
    utils.draw = mod("labelme.utils.draw", label_colormap=lambda n=256: np.zeros((n, 3)))
    mod("labelme", utils=utils, logger=types.SimpleNamespace(warn=print, info=print, warning=print))
    plt = mod("matplotlib.pyplot")
    gs = mod("matplotlib.gridspec")
    mod("matplotlib", pyplot=plt, gridspec=gs)
    from PIL import ImageFont
    real_truetype = ImageFont.truetype

    def truetype(*a, **k):
        try:
            return real_truetype(*a, **k)
        except OSError:
            return ImageFont.load_default()
    ImageFont.truetype = truetype


def import_ref(path, name):
    """Import one reference script under a private module name (two of them define the same class name)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@contextlib.contextmanager
def quiet():
    with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
        yield


# ------------------------------------------------------------------------------------------- T1
def golden_tiler(out, tmp):
    sys.path.insert(0, os.path.join(REF, "module/faster-rcnn"))
    sys.path.insert(0, os.path.join(REF, "module/espnet/test"))       # glomus_handler.py (docker/gpu.dockerfile:69-84 copies it beside)
    det = import_ref("module/faster-rcnn/detect_glomus_test.py", "ref_detect")
    for k, (sw, sh, std, mx, my, ov, power, lds) in enumerate(WC.T1_CASES):
        d = det.GlomusDetector("OPT_PAS", "unused.txt", os.path.join(tmp, "data", "site", ""), os.path.join(tmp, "o%d" % k, "a", "b"),
                               "", std, ov, 0.5)
        slide = FakeSlide(sw, sh, None, (mx, my), power, lds)
        # what split() does before scan_region (:178-186)
        d.org_slide_width, d.org_slide_height = slide.dimensions
        d.mpp_x = float(slide.properties["openslide.mpp-x"])
        d.mpp_y = float(slide.properties["openslide.mpp-y"])
        d.org_slide_objective_power = int(slide.properties["openslide.objective-power"])
        d.detect_box = lambda *a, **kw: []               # the TF detector is out of scope: no boxes
        with quiet():
            d.scan_region(None, slide, "site", "spec", "file", io.StringIO(), None, None, None, None, None)
        calls = np.array(slide.calls, np.int64).reshape(-1, 5)
        out["t1_%d_calls" % k] = calls                   # (x_start, y_start, level, window_x, window_y) per read_region
        out["t1_%d_downsample" % k] = np.float64(d.slide_downsample)
    sys.path.pop(0); sys.path.pop(0)


# ------------------------------------------------------------------------------------------- T2-T4
def png_b64(arr_u8):
    from PIL import Image
    f = io.BytesIO()
    Image.fromarray(arr_u8).save(f, format="PNG")
    return base64.b64encode(f.getvalue()).decode("utf-8")


def run_generate_pred_wsi(ev, cv2, name, tmp, black_slide):
    """Drive the real generate_pred_wsi for one case; returns what it computed."""
    c = WC.STITCH_CASES[name]
    boxes, masks = WC.stitch_inputs(name)
    key = "S_" + name + ("_b" if black_slide else "_r")
    jdir, wdir, odir = (os.path.join(tmp, key, s) for s in ("json", "wsi", "out"))
    os.makedirs(os.path.join(jdir, key)); os.makedirs(os.path.join(wdir, key)); os.makedirs(odir)
    for b, m in zip(boxes, masks):
        if c["city"]:
            m = WC.CITY_IDS[m]
        # the file name convention of make_seg_data.py:359 / VisualizeResults_iou.py:182 that overlay() searches for (:272)
        fn = "xmin{}_ymin{}_xmax{}_ymax{}.json".format(int(b[0] / 8), int(b[1] / 8), int(b[2] / 8), int(b[3] / 8))
        with open(os.path.join(jdir, key, fn), "w") as f:
            json.dump({"imageData": png_b64(m), "shapes": [], "imagePath": fn.replace("json", "PNG")}, f)
    ndpi = os.path.join(wdir, key, key + ".ndpi")
    open(ndpi, "w").close()
    rgb = np.zeros((c["sh"], c["sw"], 3), np.uint8) if black_slide else WC.slide_rgb(name)
    SLIDES[ndpi] = FakeSlide(c["sw"], c["sh"], rgb)
    g = ev.Generate_Segmentation_Gt("OPT_PAS", None, None, None, 0.01, None, odir, wdir, None, None, c["ws"], jdir, 5, False, 0, 0)
    g.detected_glomus_list[key] = [list(b) for b in boxes]
    windows, window_preds, written = [], [], {}
    real_overlay = g.overlay

    def overlay(bbox_list, times, mx, my, jl, xmin, ymin, xmax, ymax, dt):
        r = real_overlay(bbox_list, times, mx, my, jl, xmin, ymin, xmax, ymax, dt)
        windows.append((xmin, ymin, xmax, ymax))
        window_preds.append(np.array(r))
        return r
    g.overlay = overlay
    real_imwrite = cv2.imwrite
    cv2.imwrite = lambda p, a: written.__setitem__(p, np.array(a)) or True
    err = ""
    try:
        with quiet():
            g.generate_pred_wsi()
    except cv2.error as e:
        err = "cv2.error"
    finally:
        cv2.imwrite = real_imwrite
    whole = next(iter(written.values())) if written else None
    return c, boxes, masks, windows, window_preds, whole, err


def golden_stitch(out, tmp):
    import cv2
    sys.path.insert(0, os.path.join(REF, "module/common"))
    sys.path.insert(0, os.path.join(REF, "module/espnet/test"))
    ev = import_ref("module/espnet/test/eval_wsi_segmentation.py", "ref_eval_wsi")
    pal_bgr = np.array([[b, g, r] for r, g, b in ev.PALLETE[:5]], np.uint8)
    code = {tuple(cv2.addWeighted(np.zeros((1, 1, 3), np.uint8), 0.4, pal_bgr[k].reshape(1, 1, 3), 0.6, 0)[0, 0]): k for k in range(5)}
    assert len(code) == 5
    for name in WC.STITCH_CASES:
        c, boxes, masks, windows, preds, whole_b, err = run_generate_pred_wsi(ev, cv2, name, tmp, black_slide=True)
        out["s_%s_crashed" % name] = np.int64(1 if err else 0)
        out["s_%s_windows" % name] = np.array(windows, np.int64).reshape(-1, 4)      # visited by the real loop, in order
        level0 = np.zeros((c["sh"], c["sw"]), np.uint8)
        for (xmin, ymin, xmax, ymax), p in zip(windows, preds):
            assert p.shape == (ymax - ymin, xmax - xmin) and (p.size == 0 or (p.min() >= 0 and p.max() < 5))
            level0[ymin:ymax, xmin:xmax] = p
        out["s_%s_level0" % name] = level0            # every window's overlay() result at its place
        if err:
            continue                                  # reference crashed in cv2.resize (zero-sized window): nothing written
        # the /8 label image: black slide => whole = 0.6 * palette colour, decoded back to class ids
        ds8 = np.full(whole_b.shape[:2], 255, np.uint8)
        for col, k in code.items():
            ds8[(whole_b == np.array(col)).all(-1)] = k
        assert ds8.max() < 5
        out["s_%s_ds8" % name] = ds8
        _, _, _, windows_r, preds_r, whole_r, err_r = run_generate_pred_wsi(ev, cv2, name, tmp, black_slide=False)
        assert not err_r and windows_r == windows and all(np.array_equal(a, b) for a, b in zip(preds, preds_r))
        assert whole_r.min() >= 0 and whole_r.max() <= 255
        out["s_%s_render" % name] = whole_r.astype(np.uint8)       # the array handed to cv2.imwrite(<slide>_pred.jpg)
    # check_overlap
    ah = sys.modules["annotation_handler"]
    out["overlap_scores"] = np.array([ah.AnnotationHandler.check_overlap(a, b) for a, b in WC.overlap_pairs()], np.float64)
    sys.path.pop(0); sys.path.pop(0)


# ------------------------------------------------------------------------------------------- f2: crop extraction
def golden_crops(out, tmp):
    from PIL import Image
    sys.path.insert(0, os.path.join(REF, "module/common"))
    sys.path.insert(0, os.path.join(REF, "module/faster-rcnn"))
    sys.path.insert(0, os.path.join(REF, "module/espnet/test"))
    for m in ("annotation_handler", "glomus_handler", "utils", "utils.shape", "utils.my_lblsave"):
        sys.modules.pop(m, None)
    ms = import_ref("module/faster-rcnn/make_seg_data.py", "ref_make_seg")
    name = "wide"
    c = WC.STITCH_CASES[name]
    boxes, _ = WC.stitch_inputs(name)
    key = "CROPS"
    wdir, odir = os.path.join(tmp, key, "wsi"), os.path.join(tmp, key, "out")
    os.makedirs(os.path.join(wdir, key))
    ndpi = os.path.join(wdir, key, key + ".ndpi")
    open(ndpi, "w").close()
    slide = FakeSlide(c["sw"], c["sh"], WC.slide_rgb(name))
    SLIDES[ndpi] = slide
    g = ms.Generate_Segmentation_Gt("OPT_PAS", None, None, None, 0.01, odir, wdir, None, None)
    g.detected_glomus_list[key] = [list(b) for b in boxes]
    with quiet():
        g.output_org_files()
    files = sorted(os.listdir(os.path.join(odir, "org_image", key)))
    out["crop_calls"] = np.array(slide.calls, np.int64)                       # (x, y, level, w, h) per box, in box order
    names, sums = [], []
    for b in boxes:
        fn = "xmin{}_ymin{}_xmax{}_ymax{}.PNG".format(int(b[0] / 8), int(b[1] / 8), int(b[2] / 8), int(b[3] / 8))
        assert fn in files
        import cv2
        img = cv2.imread(os.path.join(odir, "org_image", key, fn))           # what VisualizeResults_iou.py:103 reads: BGR u8
        assert img.shape == (b[3] - b[1], b[2] - b[0], 3)
        names.append(fn)
        sums.append([int(img[..., ch].astype(np.int64).sum()) for ch in range(3)] + [int(img[0, 0, 0]), int(img[-1, -1, 2])])
    out["crop_names"] = np.array(names)
    out["crop_bgr_sums"] = np.array(sums, np.int64)                           # per-channel BGR sums + two probe pixels
    sys.path.pop(0); sys.path.pop(0); sys.path.pop(0)


# ------------------------------------------------------------------------------------------- evaluateModel
def golden_frontend(tmp):
    import cv2
    import torch
    from oracle import espnet_oracle as O
    sys.path.insert(0, os.path.join(REF, "module/common"))
    sys.path.insert(0, os.path.join(REF, "module/espnet/test"))
    vis = import_ref("module/espnet/test/VisualizeResults_iou.py", "ref_visualize")
    Net = sys.modules["Model"]
    torch.set_num_threads(8)
    out = {}
    for k, (ch, cw, in_w, in_h, fold, dist, seed) in enumerate(WC.FRONTEND_CASES):
        crops = O.synth_crops(dist, 2, ch, cw, seed=seed, sigma=3.0)            # BGR u8 [2,ch,cw,3]
        d = os.path.join(tmp, "F%d" % k, "rgb", "P001")
        sav = os.path.join(tmp, "F%d" % k, "out")
        os.makedirs(d); os.makedirs(sav)
        paths = []
        for i, c in enumerate(crops):
            p = os.path.join(d, "xmin%d_ymin0_xmax9_ymax9.PNG" % i)
            cv2.imwrite(p, c)                                                    # lossless; imread gives back BGR
            paths.append(p)
        mean, std = O.FOLD_MEAN_STD[fold]
        model = Net.ESPNet(5, 2, 8)
        model.load_state_dict(torch.load(os.path.join(REF, "models/espnet_fold%d.pth" % fold), map_location="cpu", weights_only=True), strict=True)
        model.eval()
        seen_in, seen_logits, seen_masks = [], [], []

        def spy_model(x):
            with torch.no_grad():
                y = model(x)
            seen_in.append(x.detach().numpy().copy())
            seen_logits.append(y.detach().numpy().copy())
            return y
        real_b2l = vis.bound2line
        vis.bound2line = lambda m, **kw: (seen_masks.append(np.array(m)), real_b2l(m, **kw))[1]
        args = types.SimpleNamespace(mean=[str(v) for v in mean], std=[str(v) for v in std], inWidth=in_w, inHeight=in_h, classes=5,
                                     savedir=sav, gpu_id=-1, modelType=1, colored=True, overlay=True, cityFormat=False, img_extn="PNG")
        import warnings
        with quiet(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vis.evaluateModel(args, spy_model, None, paths, [None] * len(paths), "cpu")
        vis.bound2line = real_b2l
        counts = [list(map(int, l.strip().split(",")[2:])) for l in open(os.path.join(sav, "summary_pixel.csv")).readlines()[1:]]
        # the crops are regenerated from (dist, seed) by the tests: O.synth_crops(dist, 2, ch, cw, seed=seed, sigma=3.0)
        out["f_%d_net_in" % k] = np.concatenate(seen_in, 0).astype(np.float32)        # what model() received (:107-123)
        if k == 1:
            out["f_%d_logits" % k] = np.concatenate(seen_logits, 0).astype(np.float32)
        out["f_%d_masks" % k] = np.stack(seen_masks).astype(np.uint8)                 # class map at crop size (:128-129)
        out["f_%d_counts" % k] = np.array(counts, np.int64)                           # summary_pixel.csv (:151-156)
    np.savez_compressed(os.path.join(HERE, "frontend_golden.npz"), **out)
    print("frontend_golden.npz", {k: v.shape for k, v in out.items()})


def main():
    install_stubs()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data", "site"))
        golden_tiler(out, tmp)
        golden_stitch(out, tmp)
        golden_crops(out, tmp)
        np.savez_compressed(os.path.join(HERE, "wsi_golden.npz"), **out)
        print("wsi_golden.npz: %d arrays, %.2f MB" % (len(out), os.path.getsize(os.path.join(HERE, "wsi_golden.npz")) / 1e6))
        golden_frontend(tmp)


if __name__ == "__main__":
    main()
