"""ctypes binding of libespnet_b200.so (include/espnet_b200.h).  The product path has no fallback:
if the library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# ESPNET_B200_LIB: developer override used to A/B kernel variants; the default is the in-tree build
LIB_PATH = os.environ.get("ESPNET_B200_LIB") or os.path.join(_HERE, "csrc", "libespnet_b200.so")

OK, EINVAL, ESHAPE, ECUDA, ESTATE, EMISSING = 0, -1, -2, -3, -4, -5
IN_F32_NCHW, IN_U8_BGR_HWC, IN_U8_SLIDE = 0, 1, 2
MODE_FP32, MODE_F16TC = 0, 1
NET_FULL, NET_ENCODER = 0, 1


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int), ("shape", C.c_int64 * 4)]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("in_fmt", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("mean", C.c_float * 3), ("std_", C.c_float * 3),
        ("origins", C.c_void_p), ("slide_h", C.c_int), ("slide_w", C.c_int),
        ("logits", C.c_void_p), ("mask", C.c_void_p), ("prob_acc", C.c_void_p),
        ("prob_init", C.c_int), ("mask_from_prob", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
    ]


# every symbol include/espnet_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "espnet_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "espnet_destroy": (None, [C.c_void_p]),
    "espnet_last_error": (C.c_char_p, [C.c_void_p]),
    "espnet_pack_weights": (C.c_int, [C.c_void_p, C.POINTER(TensorDesc), C.c_int]),
    "espnet_set_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "espnet_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "espnet_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "espnet_forward": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs)]),
    "espnet_graph_capture": (C.c_int, [C.c_void_p, C.POINTER(ForwardArgs), C.POINTER(C.c_int)]),
    "espnet_graph_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "espnet_graph_destroy": (C.c_int, [C.c_void_p, C.c_int]),
    "espnet_read_stage": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]),
    "espnet_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "espnet_get_profile": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "espnet_segment_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]),
    "espnet_stitch_boxes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "espnet_stitch_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]),
    "espnet_stitch_grid_band": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]),
    "espnet_peer_alloc": (C.c_int, [C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.c_char_p]),
    "espnet_peer_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "espnet_peer_close": (C.c_int, [C.c_void_p, C.c_int]),
    "espnet_peer_free": (C.c_int, [C.c_void_p, C.c_int]),
    "espnet_max_merge_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "espnet_ds8_lut": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "espnet_downsample_lut": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "espnet_confusion_hist": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "espnet_bilinear_lut": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "espnet_nearest_lut": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "espnet_preprocess_resize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "espnet_preprocess_resize_boxes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "espnet_resize_nearest_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "espnet_palette_overlay": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "espnet_render_ds8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "espnet_class_counts": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "espnet_tc_selftest": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "espnet_launch_count": (C.c_ulonglong, []),
    "espnet_version": (C.c_int, []),
}

_lib = None


def _up_to_date() -> bool:
    """make-style check: the library exists and is newer than every source it is built from"""
    if not os.path.isfile(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc")) if f.endswith((".cu", ".cuh", ".h", ".sh"))]
    inc = os.path.join(os.path.dirname(_HERE), "include")
    srcs += [os.path.join(inc, f) for f in os.listdir(inc)] if os.path.isdir(inc) else []
    return all(os.path.getmtime(f) <= t for f in srcs)


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU; one translation unit, ~2 min).
    Like make, nothing is recompiled when the library is newer than all of its sources (force=True or
    ESPNET_B200_REBUILD=1 recompiles anyway)."""
    if not force and not os.environ.get("ESPNET_B200_REBUILD") and _up_to_date():
        return LIB_PATH
    script = os.path.join(_HERE, "csrc", "build.sh")
    r = subprocess.run(["bash", script], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libespnet_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the library (never falls back to anything else)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("libespnet_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "-- there is no CPU / cuDNN fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def last_error(handle=None) -> str:
    s = lib().espnet_last_error(handle)
    return s.decode() if s else ""


def check(rc: int, handle=None, what: str = ""):
    if rc != OK:
        msg = last_error(handle) if handle is not None else ""
        raise RuntimeError("%s failed (code %d)%s" % (what or "libespnet_b200 call", rc, (": " + msg) if msg else ""))
