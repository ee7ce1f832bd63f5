"""Drop-in replacements for the reference's `Model.ESPNet` / `Model.ESPNet_Encoder`
(`module/espnet/test/Model.py:242-378`): same constructor arguments, same `forward(input)`,
same `state_dict` key set (so `models/espnet_fold*.pth` load with `strict=True`), but the forward
runs in libespnet_b200.so (hand-written sm_100a CUDA kernels) through the C ABI of
`include/espnet_b200.h`.  There is no PyTorch-op, cuDNN or CPU fallback: a forward on a CPU tensor,
or without the built library, raises.

The parameter tree is generated from a table (name -> shape) instead of nested layer classes: the
modules here only *hold* the tensors under the reference's names; all arithmetic lives in the kernels.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["ESPNet", "ESPNet_Encoder", "ESPNetEnsemble", "GraphedSegmenter", "FOLD_MEAN_STD"]

# per-fold BGR mean / std published by the reference (README.md:243-249)
FOLD_MEAN_STD = {
    1: ((204.60071, 170.19359, 199.57469), (20.61257, 42.92207, 28.401505)),
    2: ((202.38148, 167.13171, 198.10599), (20.704079, 42.958416, 28.366297)),
    3: ((203.12099, 167.813, 198.50894), (21.038654, 43.769535, 29.034416)),
    4: ((203.66399, 167.94217, 198.58081), (20.96783, 43.556736, 28.838718)),
    5: ((204.49896, 169.03307, 199.22058), (20.547842, 42.86628, 27.966227)),
}


# ------------------------------------------------------------------------------------------------
# parameter table
# ------------------------------------------------------------------------------------------------
def _bn_rows(key: str, c: int):
    return [(key + ".weight", (c,), "one"), (key + ".bias", (c,), "zero"),
            (key + ".running_mean", (c,), "buf_zero"), (key + ".running_var", (c,), "buf_one"),
            (key + ".num_batches_tracked", (), "buf_count")]


def _block_rows(key: str, cin: int, cout: int, down: bool):
    n = cout // 5
    n1 = cout - 4 * n
    k = 3 if down else 1
    rows = [(key + ".c1.conv.weight", (n, cin, k, k), "conv"), (key + ".d1.conv.weight", (n1, n, 3, 3), "conv")]
    rows += [(key + ".d%d.conv.weight" % d, (n, n, 3, 3), "conv") for d in (2, 4, 8, 16)]
    if down:   # DownSamplerB: .bn / .act      (Model.py:141-142)
        rows += _bn_rows(key + ".bn", cout) + [(key + ".act.weight", (cout,), "prelu")]
    else:      # ESP block: .bn is a BR         (Model.py:184)
        rows += _bn_rows(key + ".bn.bn", cout) + [(key + ".bn.act.weight", (cout,), "prelu")]
    return rows


def encoder_rows(classes: int, p: int, q: int):
    """Key set of ESPNet_Encoder(classes,p,q) in registration order (Model.py:252-271)."""
    rows = [("level1.conv.weight", (16, 3, 3, 3), "conv")] + _bn_rows("level1.bn", 16) + [("level1.act.weight", (16,), "prelu")]
    rows += _bn_rows("b1.bn", 19) + [("b1.act.weight", (19,), "prelu")]
    rows += _block_rows("level2_0", 19, 64, True)
    for i in range(p):
        rows += _block_rows("level2.%d" % i, 64, 64, False)
    rows += _bn_rows("b2.bn", 131) + [("b2.act.weight", (131,), "prelu")]
    rows += _block_rows("level3_0", 131, 128, True)
    for i in range(q):
        rows += _block_rows("level3.%d" % i, 128, 128, False)
    rows += _bn_rows("b3.bn", 256) + [("b3.act.weight", (256,), "prelu")]
    rows += [("classifier.conv.weight", (classes, 256, 1, 1), "conv")]
    return rows


def decoder_rows(classes: int):
    """Decoder-side keys of ESPNet (Model.py:330-339)."""
    c = classes
    rows = [("level3_C.conv.weight", (c, 131, 1, 1), "conv")] + _bn_rows("br", c)
    rows += [("conv.conv.weight", (c, 19 + c, 3, 3), "conv")] + _bn_rows("conv.bn", c) + [("conv.act.weight", (c,), "prelu")]
    rows += [("up_l3.0.weight", (c, c, 2, 2), "convT")]
    rows += _bn_rows("combine_l2_l3.0.bn", 2 * c) + [("combine_l2_l3.0.act.weight", (2 * c,), "prelu")]
    rows += [("combine_l2_l3.1.conv.weight", (c, 2 * c, 3, 3), "conv")] + _bn_rows("combine_l2_l3.1.bn", c)
    rows += [("combine_l2_l3.1.act.weight", (c,), "prelu")]
    rows += [("up_l2.0.weight", (c, c, 2, 2), "convT")] + _bn_rows("up_l2.1.bn", c) + [("up_l2.1.act.weight", (c,), "prelu")]
    rows += [("classifier.weight", (c, c, 2, 2), "convT")]
    return rows


class _Holder(nn.Module):
    """Anonymous container node of the parameter tree (never called)."""

    def forward(self, *a, **k):   # pragma: no cover
        raise RuntimeError("parameter holder; the arithmetic runs inside libespnet_b200.so")


def _plant(root: nn.Module, rows):
    """Create the tensors of `rows` under their dotted names with PyTorch's default initialisers
    (what the reference's nn.Conv2d / BatchNorm2d / PReLU / ConvTranspose2d constructors produce)."""
    for name, shape, kind in rows:
        parts = name.split(".")
        node = root
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _Holder())
            node = node._modules[part]
        leaf = parts[-1]
        if kind in ("conv", "convT"):
            w = torch.empty(shape)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            node.register_parameter(leaf, nn.Parameter(w))
        elif kind == "one":
            node.register_parameter(leaf, nn.Parameter(torch.ones(shape)))
        elif kind == "zero":
            node.register_parameter(leaf, nn.Parameter(torch.zeros(shape)))
        elif kind == "prelu":
            node.register_parameter(leaf, nn.Parameter(torch.full(shape, 0.25)))
        elif kind == "buf_zero":
            node.register_buffer(leaf, torch.zeros(shape))
        elif kind == "buf_one":
            node.register_buffer(leaf, torch.ones(shape))
        elif kind == "buf_count":
            node.register_buffer(leaf, torch.tensor(0, dtype=torch.long))
        else:   # pragma: no cover
            raise ValueError(kind)


# ------------------------------------------------------------------------------------------------
# engine: one C-ABI handle + packed weights + workspace cache
# ------------------------------------------------------------------------------------------------
class _Engine:
    def __init__(self, classes: int, p: int, q: int, net: int):
        self.classes, self.p, self.q, self.net = classes, p, q, net
        self.handle = None
        self.device_index = None
        self.mode = _lib.MODE_FP32
        self.options = {}
        self._ws = None

    def close(self):
        if self.handle is not None:
            _lib.lib().espnet_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ensure(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("espnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback. "
                               "Move the module and its input to a B200 with .to('cuda').")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self.handle is not None and self.device_index != idx:
            self.close()
        if self.handle is None:
            h = C.c_void_p()
            rc = _lib.lib().espnet_create(self.classes, self.p, self.q, self.net, idx, C.byref(h))
            if rc != _lib.OK:
                raise RuntimeError("espnet_create failed (code %d): %s" % (rc, _lib.last_error(None)))
            self.handle, self.device_index, self._ws = h, idx, None
            _lib.check(_lib.lib().espnet_set_mode(self.handle, self.mode), self.handle, "espnet_set_mode")
            for k, v in self.options.items():
                _lib.check(_lib.lib().espnet_set_option(self.handle, k.encode(), v), self.handle, "espnet_set_option")
        return self.handle

    def pack(self, state: Dict[str, torch.Tensor]):
        keep = []
        descs = (_lib.TensorDesc * len(state))()
        n = 0
        for name, t in state.items():
            if name.endswith("num_batches_tracked"):
                continue
            host = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(host)
            d = descs[n]
            d.name = name.encode()
            d.data = host.data_ptr()
            d.ndim = host.dim()
            for i, s in enumerate(host.shape):
                d.shape[i] = s
            n += 1
        _lib.check(_lib.lib().espnet_pack_weights(self.handle, descs, n), self.handle, "espnet_pack_weights")

    def workspace(self, B: int, H: int, W: int, device) -> torch.Tensor:
        need = _lib.lib().espnet_workspace_bytes(self.handle, B, H, W)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def forward(self, x: torch.Tensor, in_fmt: int, B: int, H: int, W: int, mean=None, std=None, origins=None,
                slide_hw=None, logits=None, mask=None, prob_acc=None, prob_init=False, mask_from_prob=False):
        a = _lib.ForwardArgs()
        a.x, a.in_fmt, a.B, a.H, a.W = x.data_ptr(), in_fmt, B, H, W
        if mean is not None:
            for i in range(3):
                a.mean[i], a.std_[i] = float(mean[i]), float(std[i])
        if origins is not None:
            a.origins, a.slide_h, a.slide_w = origins.data_ptr(), int(slide_hw[0]), int(slide_hw[1])
        a.logits = logits.data_ptr() if logits is not None else None
        a.mask = mask.data_ptr() if mask is not None else None
        a.prob_acc = prob_acc.data_ptr() if prob_acc is not None else None
        a.prob_init, a.mask_from_prob = int(prob_init), int(mask_from_prob)
        ws = self.workspace(B, H, W, x.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib().espnet_forward(self.handle, C.byref(a)), self.handle, "espnet_forward")


class _KernelBacked(nn.Module):
    """Shared machinery of the two drop-in modules."""
    _net = _lib.NET_FULL

    def _setup(self, classes: int, p: int, q: int):
        self.classes, self.p, self.q = classes, p, q
        object.__setattr__(self, "_engine", _Engine(classes, p, q, self._net))
        object.__setattr__(self, "_dirty", True)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._mark_dirty())

    def _mark_dirty(self):
        object.__setattr__(self, "_dirty", True)

    def _apply(self, fn, recurse=True):   # .to() / .cuda() / .float() re-create tensors -> repack
        r = super()._apply(fn, recurse)
        self._mark_dirty()
        return r

    def repack(self):
        """Call after editing parameters in place (load_state_dict and .to() do it automatically)."""
        self._mark_dirty()

    def set_mode(self, mode: str):
        """'fp32' (default): fp32-EQUIVALENT arithmetic, logits within 1e-3 of the reference.  With the default option
        fp32_impl=1 the block contractions run on tcgen05 as 3-term fp16 operand splits (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
        with a scaled by 1/4 and w by 4: 22-bit mantissa products, fp32 accumulate / storage); activations must stay below
        |a| < 2.6e5 (fp16 range of a/4) -- far above anything a normalised crop produces (|a| < 40 measured) -- beyond it
        the logits turn inf/NaN; set_option('fp32_impl', 0) selects plain fp32 FMA on CUDA cores (no range limit, ~4x slower).
        'f16tc': single fp16 operands on tcgen05 (fp32 accumulate / storage), held to the 0.999 mask-agreement bar only."""
        self._engine.mode = {"fp32": _lib.MODE_FP32, "f16tc": _lib.MODE_F16TC}[mode]
        if self._engine.handle is not None:
            _lib.check(_lib.lib().espnet_set_mode(self._engine.handle, self._engine.mode), self._engine.handle, "espnet_set_mode")
        return self

    def set_option(self, key: str, value: int):
        """Library tuning knob (espnet_set_option), e.g. set_option('branch_impl', 2) forces the TMA-staged kernel."""
        self._engine.options[key] = int(value)
        if self._engine.handle is not None:
            _lib.check(_lib.lib().espnet_set_option(self._engine.handle, key.encode(), int(value)), self._engine.handle, "espnet_set_option")
        return self

    def _own_state(self) -> Dict[str, torch.Tensor]:
        return self.state_dict()

    def _ready(self, device: torch.device):
        if self.training:
            raise RuntimeError("espnet_b200 is inference-only (eval-mode BatchNorm is folded into the kernels); call .eval() first")
        eng = self._engine
        fresh = eng.handle is None or eng.device_index != (device.index if device.index is not None else torch.cuda.current_device())
        eng.ensure(device)
        if fresh or self._dirty:
            eng.pack(self._own_state())
            object.__setattr__(self, "_dirty", False)
        return eng

    def _param_device(self) -> torch.device:
        return next(self.parameters()).device

    def _check_input(self, input: torch.Tensor):
        if not isinstance(input, torch.Tensor) or input.dim() != 4 or input.shape[1] != 3:
            raise RuntimeError("expected a [B,3,H,W] tensor (Model.py:341), got %s" % (tuple(input.shape) if isinstance(input, torch.Tensor) else type(input),))
        if not input.is_cuda:
            raise RuntimeError("input is on %s: espnet_b200 has no CPU fallback, move the module and the input to a CUDA device" % input.device)
        if input.device != self._param_device():
            raise RuntimeError("input is on %s but the module is on %s" % (input.device, self._param_device()))

    def read_stage(self, name: str) -> torch.Tensor:
        """Flat fp32 copy of an internal stage of the last forward (parity taps): 'b1', 'b2', 'b3', 'up_l3',
        'up_l2', 'combine_l2_l3.0', 'encoder.classifier'."""
        eng = self._engine
        n = C.c_size_t()
        _lib.check(_lib.lib().espnet_read_stage(eng.handle, name.encode(), None, 0, C.byref(n), None), eng.handle, "espnet_read_stage")
        out = torch.empty(n.value, dtype=torch.float32, device=torch.device("cuda", eng.device_index))
        st = torch.cuda.current_stream(out.device).cuda_stream
        _lib.check(_lib.lib().espnet_read_stage(eng.handle, name.encode(), out.data_ptr(), n.value, C.byref(n), st), eng.handle, "espnet_read_stage")
        return out

    def profile(self, on: bool = True):
        """Switch per-kernel CUDA-event timing on/off (clears the record)."""
        eng = self._ready(self._param_device())
        _lib.check(_lib.lib().espnet_set_profiling(eng.handle, int(on)), eng.handle, "espnet_set_profiling")

    def profile_report(self) -> Dict[str, Tuple[float, int]]:
        """{kernel name: (summed device ms, launches)} since profile(True)."""
        eng = self._engine
        names = C.create_string_buffer(64 * 64)
        ms = (C.c_float * 64)()
        cnt = (C.c_int * 64)()
        n = C.c_int()
        _lib.check(_lib.lib().espnet_get_profile(eng.handle, names, ms, cnt, 64, C.byref(n)), eng.handle, "espnet_get_profile")
        return {names.raw[64 * i:64 * (i + 1)].split(b"\0")[0].decode(): (float(ms[i]), int(cnt[i])) for i in range(n.value)}

    # -- u8 fast path: P0 normalise + forward + arg-max fused -----------------------------------------
    def segment(self, crops_u8: torch.Tensor, mean: Sequence[float], std: Sequence[float], out: Optional[torch.Tensor] = None,
                logits: Optional[torch.Tensor] = None) -> torch.Tensor:
        """crops_u8: CUDA uint8 [B,H,W,3] BGR (what cv2.imread yields, VisualizeResults_iou.py:103).
        Returns the class map uint8 [B,H,W] (VisualizeResults_iou.py:107-128 in one call)."""
        if crops_u8.dtype != torch.uint8 or crops_u8.dim() != 4 or crops_u8.shape[-1] != 3 or not crops_u8.is_cuda:
            raise RuntimeError("segment() wants a CUDA uint8 [B,H,W,3] BGR tensor")
        crops_u8 = crops_u8.contiguous()
        B, H, W, _ = crops_u8.shape
        eng = self._ready(crops_u8.device)
        if out is None:
            out = torch.empty((B, H, W), dtype=torch.uint8, device=crops_u8.device)
        eng.forward(crops_u8, _lib.IN_U8_BGR_HWC, B, H, W, mean, std, mask=out, logits=logits)
        return out

    def segment_tiles(self, slide_u8: torch.Tensor, origins: torch.Tensor, win_h: int, win_w: int, mean, std,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Tiles are read straight out of a resident slide [SH,SW,3] u8 at origins[B,2] (x0,y0) int32 -- the
        read_region of detect_glomus_test.py:272 (zero padding outside the slide) fused into the stem."""
        if slide_u8.dtype != torch.uint8 or slide_u8.dim() != 3 or slide_u8.shape[-1] != 3 or not slide_u8.is_cuda or not slide_u8.is_contiguous():
            raise RuntimeError("segment_tiles() wants a contiguous CUDA uint8 [SH,SW,3] slide")
        if origins.dtype != torch.int32 or origins.dim() != 2 or origins.shape[1] != 2 or not origins.is_contiguous() or origins.device != slide_u8.device:
            raise RuntimeError("origins must be a contiguous int32 [B,2] tensor on the slide's device")
        B = origins.shape[0]
        eng = self._ready(slide_u8.device)
        if out is None:
            out = torch.empty((B, win_h, win_w), dtype=torch.uint8, device=slide_u8.device)
        elif out.dtype != torch.uint8 or not out.is_contiguous() or out.numel() != B * win_h * win_w or out.device != slide_u8.device:
            raise RuntimeError("out must be a contiguous uint8 [B,win_h,win_w] tensor on the slide's device")
        eng.forward(slide_u8, _lib.IN_U8_SLIDE, B, win_h, win_w, mean, std, origins=origins,
                    slide_hw=(slide_u8.shape[0], slide_u8.shape[1]), mask=out)
        return out

    def capture(self, B: int, H: int, W: int, mean, std, want_logits: bool = False) -> "GraphedSegmenter":
        """Record normalise + forward + arg-max for a fixed [B,H,W] into a CUDA graph (espnet_graph_capture): one launch per
        call instead of ~30 -- the per-crop loop of VisualizeResults_iou.py:100-129 is batch 1 and launch-latency bound."""
        return GraphedSegmenter(self, B, H, W, mean, std, want_logits)

    def host_pipeline(self, B: int, H: int, W: int, mean, std, depth: int = 2) -> "HostPipeline":
        """Pipelined host-to-host segmentation of a stream of batches (see HostPipeline)."""
        return HostPipeline(self, B, H, W, mean, std, depth)

    def segment_host(self, crops_u8: np.ndarray, mean, std) -> np.ndarray:
        """Host buffers in, host masks out (the per-crop loop of VisualizeResults_iou.py:100-129, batched);
        H2D, kernels, D2H and the sync all happen inside the C call `espnet_segment_host`."""
        crops_u8 = np.ascontiguousarray(crops_u8, dtype=np.uint8)
        B, H, W, _ = crops_u8.shape
        eng = self._ready(self._param_device())
        out = np.empty((B, H, W), np.uint8)
        m = (C.c_float * 3)(*[float(v) for v in mean])
        s = (C.c_float * 3)(*[float(v) for v in std])
        _lib.check(_lib.lib().espnet_segment_host(eng.handle, crops_u8.ctypes.data, B, H, W, m, s, out.ctypes.data),
                   eng.handle, "espnet_segment_host")
        return out


class GraphedSegmenter:
    """A captured forward with its own static buffers: `run(crops_u8)` copies the crops into the static input (or use
    `.input` directly), replays the graph on the current stream and returns the static mask tensor `.mask` (and `.logits`)."""

    def __init__(self, model: "_KernelBacked", B: int, H: int, W: int, mean, std, want_logits: bool = False):
        dev = model._param_device()
        eng = model._ready(dev)
        self.model, self.eng = model, eng
        self.input = torch.zeros((B, H, W, 3), dtype=torch.uint8, device=dev)
        self.mask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        nc = model.classes
        lshape = (B, nc, H, W) if model._net == _lib.NET_FULL else (B, nc, H // 8, W // 8)
        self.logits = torch.empty(lshape, dtype=torch.float32, device=dev) if want_logits else None
        need = _lib.lib().espnet_workspace_bytes(eng.handle, B, H, W)
        self.ws = torch.empty(need, dtype=torch.uint8, device=dev)          # private: the engine's cached workspace may be re-allocated
        a = _lib.ForwardArgs()
        a.x, a.in_fmt, a.B, a.H, a.W = self.input.data_ptr(), _lib.IN_U8_BGR_HWC, B, H, W
        for i in range(3):
            a.mean[i], a.std_[i] = float(mean[i]), float(std[i])
        a.mask = self.mask.data_ptr()
        a.logits = self.logits.data_ptr() if want_logits else None
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        gid = C.c_int(-1)
        torch.cuda.synchronize(dev)
        _lib.check(_lib.lib().espnet_graph_capture(eng.handle, C.byref(a), C.byref(gid)), eng.handle, "espnet_graph_capture")
        self.graph_id = gid.value

    def run(self, crops_u8: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.model._dirty:
            raise RuntimeError("the module's weights changed after capture(); capture again")
        if crops_u8 is not None:
            self.input.copy_(crops_u8, non_blocking=True)
        st = torch.cuda.current_stream(self.input.device).cuda_stream
        _lib.check(_lib.lib().espnet_graph_launch(self.eng.handle, self.graph_id, st), self.eng.handle, "espnet_graph_launch")
        return self.mask


class HostPipeline:
    """Streaming form of the per-crop loop of VisualizeResults_iou.py:100-129 for batches that live in (pinned) HOST
    memory: `submit(crops_u8_host, masks_host)` enqueues H2D -> fused normalise + forward + arg-max -> D2H on three CUDA
    streams with `depth` device slots, so the copies of one batch overlap the kernels of its neighbours; `drain()` waits
    for everything submitted.  Host tensors must stay alive (and should be pinned) until drained."""

    def __init__(self, model, B: int, H: int, W: int, mean, std, depth: int = 2):
        # `model`: an ESPNet / ESPNet_Encoder (normalised with mean / std) or an ESPNetEnsemble (every fold has its own mean / std)
        ens = isinstance(model, ESPNetEnsemble)
        first = model.models[0] if ens else model
        dev = first._param_device()
        self.model, self.mean, self.std, self.shape, self.depth = model, mean, std, (B, H, W), depth
        if ens:
            self._prob = torch.empty((B, first.classes, H, W), dtype=torch.float32, device=dev)      # shared: one run stream
            self._run = lambda d_in, d_mask: model.segment(d_in, out=d_mask, prob=self._prob)
        else:
            self._run = lambda d_in, d_mask: model.segment(d_in, mean, std, out=d_mask)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.d_in = [torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.d_mask = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.n = 0
        for m in (model.models if ens else [model]):
            m._ready(dev)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(torch.cuda.current_stream(dev))

    def submit(self, crops_u8_host: torch.Tensor, masks_host: torch.Tensor):
        k = self.n % self.depth
        if self.n >= self.depth:
            self.s_in.wait_event(self.ev_run[k])      # slot's previous forward has consumed its input
            self.s_run.wait_event(self.ev_out[k])     # ... and its mask has left the device
        with torch.cuda.stream(self.s_in):
            self.d_in[k].copy_(crops_u8_host, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[k])
            self._run(self.d_in[k], self.d_mask[k])
            self.ev_run[k].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[k])
            masks_host.copy_(self.d_mask[k], non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.n += 1

    def drain(self):
        self.s_out.synchronize()
        self.s_run.synchronize()


class ESPNet_Encoder(_KernelBacked):
    """ESPNet-C (reference Model.py:242-304).  forward: [B,3,H,W] fp32 -> [B,classes,H/8,W/8] fp32."""
    _net = _lib.NET_ENCODER

    def __init__(self, classes=20, p=5, q=3):
        super().__init__()
        rows = encoder_rows(classes, p, q)
        _plant(self, [r for r in rows if r[0].startswith("level1.")])
        # parameter-free average-pool pyramids of the reference (Model.py:254-255); kept so that the
        # children order -- and with it ESPNet.modules[i] -- is the reference's
        self.add_module("sample1", _Holder())
        self.add_module("sample2", _Holder())
        _plant(self, [r for r in rows if not r[0].startswith("level1.")])
        self._setup(classes, p, q)

    def forward(self, input):
        self._check_input(input)
        x = input.contiguous().float()
        B, _, H, W = x.shape
        eng = self._ready(x.device)
        out = torch.empty((B, self.classes, H // 8, W // 8), dtype=torch.float32, device=x.device)
        eng.forward(x, _lib.IN_F32_NCHW, B, H, W, logits=out)
        return out

    def segment_upsampled(self, crops_u8: torch.Tensor, mean, std) -> torch.Tensor:
        """modelType 2 of the reference script: encoder, x8 bilinear up-sampling, arg-max
        (VisualizeResults_iou.py:123-128, 258-261)."""
        return self.segment(crops_u8, mean, std)


class ESPNet(_KernelBacked):
    """ESPNet = ESPNet-C encoder + light decoder (reference Model.py:306-378).
    forward: [B,3,H,W] fp32 (H, W multiples of 8) -> logits [B,classes,H,W] fp32."""
    _net = _lib.NET_FULL

    def __init__(self, classes=20, p=2, q=3, encoderFile=None):
        super().__init__()
        self.encoder = ESPNet_Encoder(classes, p, q)
        if encoderFile is not None:                       # Model.py:321-323
            self.encoder.load_state_dict(torch.load(encoderFile))
            print('Encoder loaded!')
        # the reference keeps the encoder's children in a plain list called `modules` (Model.py:325-327)
        self.modules = [m for m in self.encoder.children()]
        _plant(self, decoder_rows(classes))
        self._setup(classes, p, q)
        # loading into the encoder alone (the reference's two-stage training hand-off) must repack us too
        self.encoder.register_load_state_dict_post_hook(lambda module, incompatible: self._mark_dirty())

    def forward(self, input):
        self._check_input(input)
        x = input.contiguous().float()
        B, _, H, W = x.shape
        eng = self._ready(x.device)
        out = torch.empty((B, self.classes, H, W), dtype=torch.float32, device=x.device)
        eng.forward(x, _lib.IN_F32_NCHW, B, H, W, logits=out)
        return out


class ESPNetEnsemble:
    """Extension (BASELINE.json config 3): softmax ensemble over folds.  For fold k the
    crop is normalised with that fold's own mean/std, p_k = softmax(ESPNet_k(x_k)); the class map is
    argmax_c sum_k p_k (ties -> lowest class).  The softmax accumulation and the final arg-max are
    epilogues of the last decoder kernel."""

    def __init__(self, models: Sequence[ESPNet], mean_std: Sequence[Tuple[Sequence[float], Sequence[float]]]):
        assert len(models) == len(mean_std) and len(models) >= 1
        self.models, self.mean_std = list(models), list(mean_std)

    def host_pipeline(self, B: int, H: int, W: int, depth: int = 2) -> "HostPipeline":
        """Pipelined host-to-host ensemble segmentation of a stream of batches (see HostPipeline)."""
        return HostPipeline(self, B, H, W, None, None, depth)

    def segment(self, crops_u8: torch.Tensor, return_prob: bool = False, out: Optional[torch.Tensor] = None,
                prob: Optional[torch.Tensor] = None):
        """`out` (uint8 [B,H,W]) and `prob` (float32 [B,classes,H,W] scratch for the accumulated softmax) may be passed in
        to avoid the allocations."""
        crops_u8 = crops_u8.contiguous()
        B, H, W, _ = crops_u8.shape
        nc = self.models[0].classes
        dev = crops_u8.device
        if prob is None:
            prob = torch.empty((B, nc, H, W), dtype=torch.float32, device=dev)
        elif prob.dtype != torch.float32 or tuple(prob.shape) != (B, nc, H, W) or prob.device != dev or not prob.is_contiguous():
            raise RuntimeError("prob must be a contiguous float32 [B,classes,H,W] tensor on the crops' device")
        if out is None:
            mask = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        elif out.dtype != torch.uint8 or tuple(out.shape) != (B, H, W) or out.device != dev or not out.is_contiguous():
            raise RuntimeError("out must be a contiguous uint8 [B,H,W] tensor on the crops' device")
        else:
            mask = out
        last = len(self.models) - 1
        for k, (m, (mean, std)) in enumerate(zip(self.models, self.mean_std)):
            eng = m._ready(crops_u8.device)
            eng.forward(crops_u8, _lib.IN_U8_BGR_HWC, B, H, W, mean, std, prob_acc=prob, prob_init=(k == 0),
                        mask=mask if k == last else None, mask_from_prob=(k == last))
        if return_prob:
            return mask, prob / float(len(self.models))
        return mask
