"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the torch modules carry the reference's state_dict key set, host-side index logic matches the oracle, and the
product path fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from glomeruli_segmentation_b200 import _lib, wsi
from glomeruli_segmentation_b200.Model import ESPNet, ESPNet_Encoder
from oracle import espnet_oracle as O
from oracle import wsi_oracle as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "espnet_b200.h")).read()
    declared = set(re.findall(r"\b(espnet_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().espnet_version() >= 100


def test_library_holds_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_state_dict_key_set_matches_reference(fold_sd):
    sd = fold_sd(1)
    m = ESPNet(classes=5, p=2, q=8)
    own = m.state_dict()
    assert list(own.keys()).__len__() == 205 and set(own) == set(sd)
    for k in sd:
        assert tuple(own[k].shape) == tuple(sd[k].shape) and own[k].dtype == sd[k].dtype, k
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    e = ESPNet_Encoder(classes=5, p=2, q=8)
    e.load_state_dict(O.encoder_state_dict(sd), strict=True)
    # the reference's quirk: ESPNet.modules is a plain list of the encoder's children (Model.py:325-327)
    assert isinstance(m.modules, list) and len(m.modules) == 11
    # defaults of the reference constructors
    assert (ESPNet().classes, ESPNet().p, ESPNet().q) == (20, 2, 3)
    assert (ESPNet_Encoder().classes, ESPNet_Encoder().p, ESPNet_Encoder().q) == (20, 5, 3)


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not mounted")
def test_shipped_checkpoints_load_unchanged():
    for k in range(1, 6):
        sd = torch.load(os.path.join(REF, "models/espnet_fold%d.pth" % k), map_location="cpu", weights_only=True)
        m = ESPNet(classes=5, p=2, q=8)
        m.load_state_dict(sd, strict=True)
        assert torch.equal(m.state_dict()["encoder.level3.7.d16.conv.weight"], sd["encoder.level3.7.d16.conv.weight"])


def test_no_cpu_fallback():
    m = ESPNet(classes=5, p=2, q=8).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 16, 16))
    if not torch.cuda.is_available():
        h = C.c_void_p()
        rc = _lib.lib().espnet_create(5, 2, 8, _lib.NET_FULL, 0, C.byref(h))
        assert rc == _lib.ECUDA and b"no CPU fallback" in _lib.lib().espnet_last_error(None)
    h = C.c_void_p()
    assert _lib.lib().espnet_create(0, 2, 8, _lib.NET_FULL, 0, C.byref(h)) == _lib.ESHAPE      # classes must be in [1, 48]
    assert _lib.lib().espnet_create(49, 2, 8, _lib.NET_FULL, 0, C.byref(h)) == _lib.ESHAPE
    assert _lib.lib().espnet_create(5, 0, 8, _lib.NET_FULL, 0, C.byref(h)) == _lib.EINVAL


def test_host_tiler_matches_oracle():
    for (w, h, std, mpp, ov, ds) in [(40000, 30000, 512, 1.0, 0.1, 1.0), (40000, 30000, 512, 1.0, 0.5, 1.0),
                                     (53248, 23040, 2000, 0.2277, 0.1, 8.0), (4096, 3072, 512, 1.0, 0.1, 1.0),
                                     (1000, 700, 96, 0.93, 0.25, 1.0)]:
        o, nx, ny, wx, wy, sx, sy = W.tile_grid(w, h, std, mpp, mpp, ov, ds)
        g = wsi.tile_grid(w, h, std, mpp, mpp, ov, ds)
        assert (g.n_x, g.n_y, g.win_x, g.win_y, g.stride_x, g.stride_y) == (nx, ny, wx, wy, sx, sy)
        assert np.array_equal(g.origins().astype(np.int64), o)
        r0, rows = wsi.shard_rows(ny, 1, 3)
        assert np.array_equal(g.origins(r0, rows).astype(np.int64), o[r0 * nx:(r0 + rows) * nx])


def test_shard_rows_partition():
    for n in (1, 7, 66, 118):
        for world in (1, 2, 4, 8):
            bands = [wsi.shard_rows(n, r, world) for r in range(world)]
            assert bands[0][0] == 0 and sum(b[1] for b in bands) == n
            for a, b in zip(bands, bands[1:]):
                assert a[0] + a[1] == b[0]


def test_ds8_lut_host_matches_oracle():
    # espnet_ds8_lut is pure host arithmetic, so it can be checked without a GPU
    for (w, h, ws) in [(40000, 30000, 2400), (6656, 2880, 2400), (1000, 760, 240), (1000, 3000, 400), (999, 700, 160)]:
        ys, xs = wsi.ds8_luts(w, h, ws)
        lvl = (np.arange(h)[:, None] * 7 + np.arange(w)[None, :] * 3) % 251
        exp = np.zeros((int(h / 8), int(w / 8)), np.int64)
        for (xmin, ymin, xmax, ymax) in W.stitch_windows(w, h, ws):
            if xmax == xmin or ymax == ymin:
                continue
            small = W.resize_nearest(lvl[ymin:ymax, xmin:xmax], int((xmax - xmin) / 8), int((ymax - ymin) / 8))
            exp[ymin // 8:ymax // 8, xmin // 8:xmax // 8] = small
        got = np.where((ys[:, None] >= 0) & (xs[None, :] >= 0), lvl[np.maximum(ys, 0)][:, np.maximum(xs, 0)], 0)
        assert np.array_equal(got, exp), (w, h, ws)
        assert wsi.stitch_y_limit(w, h, ws) == max([b[3] for b in W.stitch_windows(w, h, ws)] + [0])
    with pytest.raises(RuntimeError):
        wsi.ds8_luts(1000, 700, 100)
