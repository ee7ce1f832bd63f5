#!/usr/bin/env python
"""bench.py -- ESPNet inference throughput on B200 (BASELINE.json metric: ESPNet 512x512 crops/s & WSI Mpx/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

The headline line is BASELINE configs[1] (ESPNet-C encoder-only, batch 64 x 512x512, fp32-equivalent, per GPU).
A step = one pass of the hot path over one batch.  `value` is timed with the (normalised fp32) batch resident in HBM;
`e2e` goes through the public u8 API with pinned HOST crops in and HOST masks out, the H2D / D2H copies inside the timed
region.  One process per GPU (torchrun for N > 1); no collective in the forward (weak scaling: every rank has its own batch).

With no --workload the same JSON line also carries, under "configs", one record per remaining BASELINE config:
  espnet_b64_fp32          full ESPNet (north_star's target), batch 64, logits + arg-max
  espnet_b256_ens5_f16tc   configs[2]: full ESPNet, batch 256, folds 1-5 softmax ensemble, reduced-precision mode
  batch1_latency           configs[0]: one 512x512 crop, forward + arg-max, CUDA-graph replay, next to the CPU port at batch 1
  wsi_40000x30000          configs[3]: overlapping-tile slide, tile rows sharded over the N GPUs (strong scaling), band gather timed apart
  sweep                    configs[4]: crop 256/512/1024 x batch 1..1024 (GPU, this N) and the CPU port at small batches (N = 1)
  sustained                the headline workload looped for >= 2 s with the median SM clock
--workload X runs that single workload as the headline (wsi prints the WSI line).

--impl reference: the reference's CPU implementation of the same path.  The reference is pure Python on PyTorch and cannot
travel to the GPU box, so this arm runs the oracle port (oracle/espnet_oracle.py, pinned to the real reference through
tests/golden) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = None


def claim_stdout():
    """stdout carries the JSON line and nothing else: whatever libraries write to fd 1 (NCCL prints its version there) goes to
    stderr from here on.  Called by main() only -- importing this module must not touch the file descriptors."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

HBM_FALLBACK_GBS = 6650.0           # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # nominal CUDA-core fp32 roof
FLOP_PER_CROP_FULL = 3.636e9
FLOP_PER_CROP_ENC = 3.636e9 - (154.7 + 18.0 + 21.5) * 1e6   # minus S8..S10 (decoder-only stages)
DTYPE = {"fp32": "f16x3-split on tcgen05 (~22-bit products), f32 accumulate / storage (fp32-equivalent, 1e-3 logit bar)",
         "f16tc": "f16 operands on tcgen05, f32 accumulate / storage (0.999 mask-agreement bar)"}


def load_weights(fold, encoder_only):
    path = os.path.join(ROOT, "tests", "golden", "weights_fold%d.npz" % fold)
    if not os.path.isfile(path):
        return None
    z = np.load(path)
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    if encoder_only:
        sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    return sd


def synth_u8(B, H, W, seed):
    """D1 of SURVEY.md 8(d): iid uniform BGR u8 crops (the hardest case for low-precision paths)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.begin = index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self, timeout=5.0):
        """Call right before a timed region: waits until nvidia-smi delivers (its start-up takes longer than a short
        timed region) and discards what was sampled before."""
        if self.proc is None:
            return
        t0 = time.time()
        while not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)
        self.begin = len(self.lines)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in (self.lines[self.begin:] or self.lines):
            f = [t.strip() for t in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        return self.summary()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def pin_to_gpu_numa(local):
    """Bind this rank to the CPU cores NVML reports as local to its GPU BEFORE any pinned buffer is allocated (first-touch
    places pinned pages on that NUMA node).  Returns a description for the JSON line."""
    info = {"cpus_before": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        cpus = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(pinned=True, cpus=len(cpus), first_cpu=cpus[0], last_cpu=cpus[-1])
        else:
            info.update(pinned=False, why="empty NVML affinity")
    except Exception as e:      # no NVML / not permitted: run unpinned and say so
        info.update(pinned=False, why=str(e)[:80])
    return info


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------------
def cpu_forward_fn(workload, sample, crop=512):
    from oracle import espnet_oracle as O       # checker / CPU baseline only -- never on the product path
    enc = workload == "espnet_c_b64_fp32"
    sd = load_weights(1, False)
    if sd is None:
        sd = O.random_state_dict(5, 2, 8, seed=0)
    mean, std = O.FOLD_MEAN_STD[1]
    u8 = synth_u8(sample, crop, crop, 1234)
    esd = O.encoder_state_dict(sd)
    sds = None
    if workload == "espnet_b256_ens5":
        sds = [load_weights(k, False) or O.random_state_dict(5, 2, 8, seed=k) for k in range(1, 6)]

    def step():
        if sds is not None:
            return O.ensemble_mask(sds, u8, range(1, 6))[0]
        x = torch.from_numpy(O.normalise_bgr_u8(u8, mean, std))
        if enc:
            return O.argmax_mask(O.upsample8_bilinear(O.espnet_encoder_forward(esd, x)))
        return O.argmax_mask(O.espnet_forward(sd, x))
    return step


def time_cpu(workload, sample, steps, warmup, crop=512):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_forward_fn(workload, sample, crop)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload or "espnet_c_b64_fp32"
    sample = 8 if wl != "espnet_b256_ens5" else 2
    val, dt = time_cpu(wl, sample, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "ESPNet 512x512 crops/s", "value": val, "unit": "crops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, sample),
        "cpu_baseline": {"value": val, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": "%d synthetic 512x512 crops per step, oracle port of Model.py on torch %s CPU" % (sample, torch.__version__)},
        "e2e": {"value": val, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(workload, batch):
    names = {
        "espnet_c_b64_fp32": "ESPNet-C encoder-only (classes=5,p=2,q=8), %d synthetic 512x512 BGR crops per GPU per step, fp32" % batch,
        "espnet_b64_fp32": "full ESPNet (classes=5,p=2,q=8), %d synthetic 512x512 crops per GPU per step, fp32 logits + arg-max" % batch,
        "espnet_b256_ens5": "full ESPNet (5,2,8), %d synthetic 512x512 crops per GPU per step, folds 1-5 softmax ensemble" % batch,
    }
    return {"workload": names[workload], "crop": "512x512", "batch_per_gpu": batch, "weights": "espnet_fold1 (tests/golden)",
            "l2": "inputs (201 MB fp32 / 50 MB u8 per batch of 64) and activations (>2 GB) exceed the 126 MB L2; no flush needed",
            "timing": "`value` and `e2e` are separate legs: each starts from an idle GPU (1 s pause before the e2e leg), runs W warm-up steps and "
                      "times K steps with CUDA events; `configs.sustained` is the same workload looped for >= 2.5 s under the power cap"}


def kernel_roofline(name, per_launch_ms, B, hbm_peak, peak_src):
    """HBM roofline of ONE launch of an ESP branch-stage kernel at 512x512 crops.  `achieved` uses SURVEY.md 8(d)'s
    algorithmic bytes: the block's input read once + its output written once, fp32 (S3 / S6: 1.049 / 0.524 Melem each way
    per crop); the reduced map o1 is produced and consumed inside the fused unit and is not counted.  `with_o1` adds the
    o1 read (fp16, L2-resident between the reduce and the branch kernel) that this two-kernel implementation does perform."""
    P4, P8 = 128 * 128, 64 * 64
    l3 = name.endswith("_l3")
    if not name.startswith("esp_branch"):
        return None
    alg = B * ((128 + 128) * P8 if l3 else (64 + 64) * P4) * 4
    if "_tc3_" in name:
        o1 = B * (2 * 32 * 2 * P8 if l3 else 2 * 16 * 2 * P4)
    elif "_tc_" in name:
        o1 = B * (32 * 2 * P8 if l3 else 16 * 2 * P4)
    else:
        o1 = B * (25 * P8 if l3 else 12 * P4) * 4
    flops = B * (P8 * 2 * 9 * 25 * 128 if l3 else P4 * 2 * 9 * 12 * 64)
    ach = alg / (per_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for tp in ("r02_traffic.json", "r01_traffic.json"):     # DRAM bytes per launch from the committed ncu --set full captures
        tp = os.path.join(ROOT, "profiles", tp)
        if B == 64 and os.path.isfile(tp):
            t = json.load(open(tp)).get(name)
            if t:
                traffic, traffic_src = t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
                break
    roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launch_ms": per_launch_ms,
            "algorithmic_bytes_per_launch": alg, "algorithmic_bytes_source": "SURVEY.md 8(d) S3/S6: block input + output, fp32",
            "with_o1": {"bytes_per_launch": alg + o1, "achieved": (alg + o1) / (per_launch_ms * 1e-3) / 1e9,
                        "frac": (alg + o1) / (per_launch_ms * 1e-3) / 1e9 / hbm_peak}}
    tf = flops / (per_launch_ms * 1e-3) / 1e12
    if "_tc3_" in name:
        roof["note"] = ("fp32-equivalent mode: the contraction runs on tcgen05 as 3 fp16 MMAs per product term set (3-term operand splits); "
                        "the kernel's floor is the fixed ~57-cycle cost of small-N tcgen05.mma (DESIGN.md 4.1), the HBM roof is what SURVEY.md 8(d) "
                        "prescribes as denominator; tensor TFLOP/s below count the useful fp32-equivalent FLOPs once")
        roof["tensor"] = {"achieved_tflops_useful": tf, "mma_flops_issued_factor": 3 * (32 * 32) / (25.0 * 26.6)}
    elif "_tc_" in name:
        roof["note"] = "tcgen05 mode: the contraction runs on tensor cores, the kernel is bounded by HBM (SURVEY.md 8(d))"
        roof["tensor"] = {"achieved_tflops_useful": tf}
    else:
        roof["note"] = "fp32 mode: this kernel is CUDA-core FMA bound, not HBM bound (SURVEY.md 8(d)); see fp32_fma"
        roof["fp32_fma"] = {"achieved_tflops": tf, "peak_tflops": FP32_FMA_PEAK_TFLOPS, "frac": tf / FP32_FMA_PEAK_TFLOPS}
    return roof


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of one rank."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
        self.host = pin_to_gpu_numa(self.local)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, join=None):
        """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        if join is not None:
            join(True)          # side streams start after e0
        for _ in range(steps):
            fn()
        if join is not None:
            join(False)         # current stream waits for the side streams before e1
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        self.barrier()
        return self.max_over_ranks(ms)

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = torch.tensor([float(v)], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def make_model(ctx, encoder, fold, mode, fp32_impl="auto"):
    from glomeruli_segmentation_b200 import ESPNet, ESPNet_Encoder
    model = ESPNet_Encoder(5, 2, 8) if encoder else ESPNet(5, 2, 8)
    sd = load_weights(fold, encoder)
    if sd is not None:
        model.load_state_dict(sd, strict=True)
    model = model.to(ctx.dev).eval().set_mode(mode)
    if fp32_impl != "auto":
        model.set_option("fp32_impl", 1 if fp32_impl == "tc3" else 0)
    return model


def join_fn(ctx, pipe):
    def join(start):
        cur = torch.cuda.current_stream(ctx.dev)
        for st in (pipe.s_in, pipe.s_run, pipe.s_out):
            (st.wait_stream(cur) if start else cur.wait_stream(st))
    return join


def profile_kernels(model, step, n):
    model.profile(True)
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    rep = model.profile_report()
    model.profile(False)
    tot = sum(v[0] for v in rep.values()) or 1.0
    return rep, {k: {"ms_per_step": v[0] / n, "launches_per_step": v[1] / n, "share": v[0] / tot} for k, v in rep.items()}


def measure_crops(ctx, args, wl, mode, B, steps, warmup, sampler=None, want_other=False, want_profile=True):
    """One crops workload on this rank's GPU: resident-input throughput, end-to-end throughput from pinned host memory,
    per-kernel profile.  Returns a dict (rank-independent numbers are MAX-over-ranks times)."""
    from glomeruli_segmentation_b200 import ESPNetEnsemble, FOLD_MEAN_STD, _lib
    H = W = 512
    mean, std = FOLD_MEAN_STD[1]
    enc = wl == "espnet_c_b64_fp32"
    model = make_model(ctx, enc, 1, mode, args.fp32_impl)
    ens = None
    if wl == "espnet_b256_ens5":
        models = [model] + [make_model(ctx, False, k, mode, args.fp32_impl) for k in range(2, 6)]
        ens = ESPNetEnsemble(models, [FOLD_MEAN_STD[k] for k in range(1, 6)])
    u8_host = torch.from_numpy(synth_u8(B, H, W, 1234 + ctx.rank)).pin_memory()
    mask_host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    u8_dev = u8_host.to(ctx.dev)
    mask_dev = torch.empty((B, H, W), dtype=torch.uint8, device=ctx.dev)
    logits = None
    if ens is not None:
        def step_resident():
            return ens.segment(u8_dev)
    else:
        # device-resident normalised fp32 batch = what the reference's forward receives (P0 done once, outside the timing)
        m_t = torch.tensor(mean, device=ctx.dev, dtype=torch.float32)
        s_t = torch.tensor(std, device=ctx.dev, dtype=torch.float32)
        x_dev = (((u8_dev.float() - m_t) / s_t) / 255.0).permute(0, 3, 1, 2).contiguous()
        if enc:
            def step_resident():
                return model(x_dev)
        else:
            logits = torch.empty((B, 5, H, W), dtype=torch.float32, device=ctx.dev)

            def step_resident():
                model._engine.forward(x_dev, _lib.IN_F32_NCHW, B, H, W, logits=logits, mask=mask_dev)
                return logits
    pipe = model.host_pipeline(B, H, W, mean, std, depth=args.depth) if ens is None else ens.host_pipeline(B, H, W, depth=2)

    def step_e2e():
        if pipe is not None:     # public streaming API: H2D / kernels / D2H of consecutive batches overlap
            pipe.submit(u8_host, mask_host)
            return
        d = u8_host.to(ctx.dev, non_blocking=True)
        mask_host.copy_(ens.segment(d), non_blocking=True)

    for _ in range(warmup):
        step_resident()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.mark()
        step_resident()         # the GPU idled while nvidia-smi started up
    l0 = _lib.lib().espnet_launch_count()
    ms = ctx.timed(step_resident, steps)
    launches = int(_lib.lib().espnet_launch_count() - l0)
    # per-kernel CUDA events on the same stream, straight after the timed steps (same thermal state as `value`)
    rep, kernels = (None, None)
    if want_profile:
        rep, kernels = profile_kernels(model, step_resident, min(steps, 5))
    torch.cuda.synchronize()
    time.sleep(1.0)             # the e2e leg starts like the resident leg did: idle GPU, then W warm-up steps, then K timed steps
    for _ in range(warmup):
        step_e2e()
    if pipe is not None:
        pipe.drain()
    ms_e2e = ctx.timed(step_e2e, steps, join_fn(ctx, pipe) if pipe is not None else None)
    out = {"value": ctx.world * B * steps / (ms * 1e-3), "unit": "crops/s", "ms_per_step": ms / steps, "batch_per_gpu": B, "mode": mode,
           "gpu_launches": launches,
           "e2e": {"value": ctx.world * B * steps / (ms_e2e * 1e-3), "unit": "crops/s", "h2d_bytes_per_step": int(u8_host.numel()),
                   "d2h_bytes_per_step": int(mask_host.numel()), "ms_per_step": ms_e2e / steps,
                   "api": ("model.host_pipeline(depth=%d).submit(pinned host u8 crops, pinned host u8 masks): H2D, fused normalise + forward + arg-max, "
                           "D2H every step on 3 streams" % args.depth) if ens is None else
                          "ESPNetEnsemble.host_pipeline(depth=2).submit(pinned host u8 crops, pinned host u8 masks): H2D, 5 forwards with softmax "
                          "accumulate + arg-max, D2H every step on 3 streams"}}
    if want_profile:
        out["kernels"] = kernels
        out["top_kernel"] = max(rep, key=lambda k: rep[k][0]) if rep else None
    if ens is not None:
        # per-fold mask agreement of the reduced-precision mode with the fp32-equivalent mode on this batch (D1: the hard input)
        sub = u8_dev[:32]
        agree = []
        for k, mk in enumerate(ens.models, start=1):
            mk.set_mode(mode)
            a = mk.segment(sub, *FOLD_MEAN_STD[k]).clone()
            mk.set_mode("fp32")
            b = mk.segment(sub, *FOLD_MEAN_STD[k])
            mk.set_mode(mode)
            agree.append(float((a == b).float().mean().item()))
        m_mode = ens.segment(sub).clone()
        for mk in ens.models:
            mk.set_mode("fp32")
        m_fp32 = ens.segment(sub)
        for mk in ens.models:
            mk.set_mode(mode)
        out["per_fold_mask_agreement_with_fp32_mode"] = agree
        out["ensemble_mask_agreement_with_fp32_mode"] = float((m_mode == m_fp32).float().mean().item())
    if want_other and ens is None:
        # second leg: the other compute mode on the same resident batch (reported beside the headline, never instead of it)
        om = "f16tc" if mode == "fp32" else "fp32"
        ref_mask = model.segment(u8_dev, mean, std).clone()
        model.set_mode(om)
        for _ in range(3):
            step_resident()
        ms_o = ctx.timed(step_resident, steps)
        rep_o, k_o = profile_kernels(model, step_resident, min(steps, 5))
        agree = float((model.segment(u8_dev, mean, std) == ref_mask).float().mean().item())
        model.set_mode(mode)
        out["other_mode"] = {"mode": om, "value": ctx.world * B * steps / (ms_o * 1e-3), "unit": "crops/s", "ms_per_step": ms_o / steps,
                             "mask_agreement_with_%s" % mode: agree, "kernels": k_o,
                             "top_kernel": max(rep_o, key=lambda k: rep_o[k][0]) if rep_o else None}
    out["_model"], out["_step"], out["_rep"] = model, step_resident, rep
    return out


def add_rooflines(rec, B):
    hbm_peak, peak_src = peaks()
    for holder in (rec, rec.get("other_mode") or {}):
        for kname, kk in (holder.get("kernels") or {}).items():
            r = kernel_roofline(kname, kk["ms_per_step"] / kk["launches_per_step"], B, hbm_peak, peak_src)
            if r is not None:
                holder.setdefault("rooflines", {})[kname] = r
    top = rec.get("top_kernel")
    rec["roofline"] = (rec.get("rooflines") or {}).get(top)
    # the fused unit SURVEY.md 8(d) defines (S6: one level-3 ESP block = reduce + branch): block bytes over both kernels' time
    k = rec.get("kernels") or {}
    for br, rd in (("esp_branch_tc3_l3", "reduce1x1_tc3_l3"), ("esp_branch_tc_l3", "reduce1x1_tc_l3")):
        if br in k and rd in k:
            t_blk = (k[rd]["ms_per_step"] / k[rd]["launches_per_step"] + k[br]["ms_per_step"] / k[br]["launches_per_step"]) * 1e-3
            bytes_blk = B * (128 + 128) * 64 * 64 * 4
            rec["block_roofline_l3"] = {"unit": "reduce1x1 + esp_branch (one level-3 ESP block, Model.py:187-214)", "bytes": bytes_blk,
                                        "time_ms": t_blk * 1e3, "achieved": bytes_blk / t_blk / 1e9, "peak": hbm_peak, "frac": bytes_blk / t_blk / 1e9 / hbm_peak}


def strip_private(rec):
    return {k: v for k, v in rec.items() if not k.startswith("_")}


def measure_batch1(ctx, args):
    """BASELINE configs[0]: ESPNet(5,2,8) fold1, ONE 512x512 crop, forward + arg-max.  Device-timed latency of the plain
    forward (29 launches) and of the CUDA-graph replay (1 launch), host-to-host latency through the graph, CPU port beside."""
    from glomeruli_segmentation_b200 import FOLD_MEAN_STD
    mean, std = FOLD_MEAN_STD[1]
    model = make_model(ctx, False, 1, "fp32", args.fp32_impl)
    u8_host = torch.from_numpy(synth_u8(1, 512, 512, 99 + ctx.rank)).pin_memory()
    mask_host = torch.empty((1, 512, 512), dtype=torch.uint8).pin_memory()
    u8 = u8_host.to(ctx.dev)
    out = torch.empty((1, 512, 512), dtype=torch.uint8, device=ctx.dev)
    g = model.capture(1, 512, 512, mean, std)
    g.input.copy_(u8)
    n = 200

    def plain():
        model.segment(u8, mean, std, out=out)

    def graph():
        g.run()

    def host_to_host():
        g.input.copy_(u8_host, non_blocking=True)
        g.run()
        mask_host.copy_(g.mask, non_blocking=True)
        torch.cuda.current_stream(ctx.dev).synchronize()          # a per-crop loop consumes each mask before the next crop

    res = {}
    for name, fn, k in (("plain_forward", plain, n), ("graph_replay", graph, n)):
        for _ in range(10):
            fn()
        res[name + "_ms"] = ctx.timed(fn, k) / k
    for _ in range(10):
        host_to_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        host_to_host()
    res["host_to_host_ms"] = ctx.max_over_ranks((time.perf_counter() - t0) / n * 1e3)
    assert torch.equal(g.mask, model.segment(u8, mean, std))
    res.update(unit="ms per 512x512 crop (batch 1)", crops_per_s_graph=1e3 / res["graph_replay_ms"],
               note="plain = 29 kernel launches per crop (programmatically chained); graph = espnet_graph_launch; host_to_host = pinned H2D + graph + D2H + stream sync per crop (wall clock)")
    if ctx.world == 1 and not args.no_cpu:
        v, dt = time_cpu("espnet_b64_fp32", 1, 5, 2)
        res["cpu_port_ms"] = dt * 1e3
        res["cpu_cores"] = torch.get_num_threads()
    return res


def measure_sustained(ctx, rec, seconds=2.5):
    """The headline workload back to back for >= `seconds` with the clocks sampled over that region only."""
    step, B = rec["_step"], rec["batch_per_gpu"]
    per = max(rec["ms_per_step"], 1e-3)
    steps = int(seconds * 1e3 / per) + 1
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
        sampler.mark()
    ms = ctx.timed(step, steps)
    clocks = sampler.stop() if ctx.rank == 0 else None
    return {"seconds": ms * 1e-3, "steps": steps, "value": ctx.world * B * steps / (ms * 1e-3), "unit": "crops/s", "ms_per_step": ms / steps, "clocks": clocks}


def measure_sweep(ctx, args):
    """BASELINE configs[4]: full ESPNet fold1 (fp32-equivalent mode), device-resident u8 crops in, u8 masks out
    (normalise + forward + arg-max fused), crop 256/512/1024 x batch 1..1024 on this run's N GPUs; CPU port at small batches."""
    from glomeruli_segmentation_b200 import FOLD_MEAN_STD
    mean, std = FOLD_MEAN_STD[1]
    model = make_model(ctx, False, 1, "fp32", args.fp32_impl)
    pts = []
    for crop in (256, 512, 1024):
        for B in (1, 4, 16, 64, 256, 1024):
            if B * crop * crop > (1 << 28):          # workspace ~130 B per input pixel: cap a point at ~35 GB
                continue
            u8 = torch.randint(0, 256, (B, crop, crop, 3), dtype=torch.uint8, device=ctx.dev)
            out = torch.empty((B, crop, crop), dtype=torch.uint8, device=ctx.dev)

            def step():
                model.segment(u8, mean, std, out=out)
            for _ in range(2):
                step()
            k = max(3, min(50, int(2e7 / (B * crop * crop)) + 1))
            ms = ctx.timed(step, k)
            cps = ctx.world * B * k / (ms * 1e-3)
            pts.append({"crop": crop, "batch_per_gpu": B, "crops_per_s": cps, "mpx_per_s": cps * crop * crop / 1e6, "ms_per_step": ms / k})
            del u8, out
        model._engine._ws = None
        torch.cuda.empty_cache()
    res = {"gpu": pts, "n_gpus": ctx.world, "mode": "fp32", "note": "points with batch x crop^2 > 2^28 px are skipped (1024^2 stops at batch 256)"}
    if ctx.world == 1 and not args.no_cpu:
        cpu = []
        for crop, B, k in ((256, 1, 3), (256, 8, 2), (512, 1, 3), (512, 8, 2), (1024, 1, 2), (1024, 4, 1)):
            v, dt = time_cpu("espnet_b64_fp32", B, k, 1, crop)
            cpu.append({"crop": crop, "batch": B, "crops_per_s": v, "mpx_per_s": v * crop * crop / 1e6, "ms_per_step": dt * 1e3})
        res["cpu_port"] = {"cores": torch.get_num_threads(), "points": cpu}
    return res


def synth_slide_rows(dev, sw, y0, y1, seed=1234, block=2048):
    """Rows [y0, y1) of the synthetic slide: every 2048-row block has its own seeded generator, so any rank can build just
    the band it needs and all ranks agree on the pixels."""
    out = torch.empty((y1 - y0, sw, 3), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    for b in range(y0 // block, (y1 + block - 1) // block):
        g.manual_seed(seed + b)
        rows = torch.randint(0, 256, (block, sw, 3), generator=g, device=dev, dtype=torch.uint8)
        lo, hi = max(y0, b * block), min(y1, (b + 1) * block)
        out[lo - y0:hi - y0] = rows[lo - b * block:hi - b * block]
        del rows
    return out


def measure_wsi(ctx, args, steps, warmup):
    """BASELINE configs[3]: synthetic slide, T1 tiler (512 px windows, overlap), tile rows sharded across ranks; every rank
    holds ONLY the slide rows its tiles read and stitches ONLY its band; band gather to rank 0 (point-to-point, overlap strips
    max-merged), T4 /8 mask.  value = slide megapixels per second for the whole job (strong scaling: the slide is fixed)."""
    from glomeruli_segmentation_b200 import FOLD_MEAN_STD, _lib, wsi
    sw, sh = (int(v) for v in args.slide.lower().split("x"))
    mean, std = FOLD_MEAN_STD[1]
    model = make_model(ctx, False, 1, args.mode, args.fp32_impl)
    grid = wsi.tile_grid(sw, sh, 512, 1.0, 1.0, args.overlap, 1.0)
    k0, k1, y0, y1 = wsi.band_tiles(grid, sh, ctx.rank, ctx.world)
    slide = synth_slide_rows(ctx.dev, sw, y0, y1)
    batch = args.batch or 256
    tm = {}
    gather, gather_note = None, "single GPU"
    if ctx.world > 1:
        if args.gather == "p2p":
            try:        # collective; raises on EVERY rank or on none
                gather = wsi.PeerGather(grid, sh, sw, ctx.rank, ctx.world, ctx.dev)
                gather_note = "p2p: stitch kernels write the bands into rank 0's slide mask through CUDA-IPC mappings (NVLink), overlap strips max-merged"
            except RuntimeError as e:
                gather_note = "nccl send/recv (peer mapping unavailable: %s)" % str(e)[:120]
        else:
            gather_note = "nccl send/recv of the band masks, overlap strips max-merged"

    def step():
        return wsi.segment_slide(model, slide, mean, std, std_size=512, mpp=1.0, overlap=args.overlap, batch=batch, rank=ctx.rank,
                                 world=ctx.world, reduce_to_rank0=True, slide_y0=y0, slide_h=sh, timings=tm, gather=gather)
    for _ in range(warmup):
        step()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    l0 = _lib.lib().espnet_launch_count()
    e0.record()
    for _ in range(steps):
        level0, ds8, n_local = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = int(_lib.lib().espnet_launch_count() - l0)
    ctx.barrier()
    ms = ctx.max_over_ranks(ms)
    fwd = ctx.max_over_ranks(tm.get("forward_ms", 0.0))
    gather_ms = ctx.max_over_ranks(tm.get("gather_ms", 0.0)) if ctx.world > 1 else 0.0
    place_ms = ctx.max_over_ranks(tm.get("place_kernels_ms", 0.0)) if ctx.world > 1 else 0.0
    clocks = sampler.stop() if ctx.rank == 0 else None
    hist = torch.bincount(ds8.reshape(-1).long(), minlength=5).tolist() if ctx.rank == 0 else None
    if gather is not None:
        del level0, ds8
        gather.close()
    if ctx.rank != 0:
        return None
    mpx = sw * sh / 1e6
    return {
        "metric": "WSI Mpx/s", "value": mpx * steps / (ms * 1e-3), "unit": "Mpx/s", "n_gpus": ctx.world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": DTYPE[args.mode], "data": "synthetic",
        "config": {"workload": "synthetic %dx%d px BGR u8 slide (each rank holds the rows of its band), T1 tiler 512 px / overlap %.2f -> %d tiles (%dx%d), "
                               "full ESPNet(5,2,8) fold1 + arg-max per tile, per-rank band stitch (T3) placed on rank 0, "
                               "overlap strips max-merged, T4 /8 mask" % (sw, sh, args.overlap, grid.count, grid.n_x, grid.n_y),
                   "tile_batch": batch, "mode": args.mode, "l2": "slide band (%.1f GB) and tile activations exceed the 126 MB L2" % ((y1 - y0) * sw * 3 / 1e9)},
        "tiles_per_s": grid.count * steps / (ms * 1e-3), "tile_mpx_per_s": grid.count * 0.262144 * steps / (ms * 1e-3),
        "phases_ms_max_over_ranks": {"tiles_forward": fwd, "band_gather": gather_ms, "band_gather_share": gather_ms / (ms / steps) if ms else None,
                                     "band_place_kernels": place_ms,
                                     "note": "band_gather = from the end of a rank's own tiles to the assembled mask on rank 0, i.e. it contains the wait for the "
                                             "slowest rank and two barriers; band_place_kernels = the stitch kernels that write over NVLink, alone (p2p)"},
        "gather": gather_note, "gather_bytes_received_rank0": tm.get("bytes_received"), "gather_bytes_max_merged": tm.get("bytes_merged"),
        "gpu_launches": launches, "clocks": clocks, "ds8_class_histogram": hist,
    }


def free_rec(rec):
    for k in ("_model", "_step", "_rep"):
        rec.pop(k, None)
    torch.cuda.empty_cache()


def run_ours(args):
    ctx = Ctx()
    if args.workload == "wsi":
        line = measure_wsi(ctx, args, args.steps, max(args.warmup, 1))
        if ctx.rank == 0:
            emit(line)
        ctx.close()
        return
    wl = args.workload or "espnet_c_b64_fp32"
    all_configs = args.workload is None and not args.quick
    B = args.batch or {"espnet_c_b64_fp32": 64, "espnet_b64_fp32": 64, "espnet_b256_ens5": 256}[wl]
    warm = max(args.warmup, 3)
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
    head = measure_crops(ctx, args, wl, args.mode, B, args.steps, warm, sampler if ctx.rank == 0 else None, want_other=not args.single_mode)
    clocks = sampler.stop() if ctx.rank == 0 else None      # sampled over both timed regions (resident + end-to-end)
    add_rooflines(head, B)
    extra = {}
    if all_configs:
        extra["sustained"] = measure_sustained(ctx, head)
    free_rec(head)
    if all_configs:
        rec = measure_crops(ctx, args, "espnet_b64_fp32", "fp32", 64, args.steps, warm)
        add_rooflines(rec, 64)
        free_rec(rec)
        extra["espnet_b64_fp32"] = dict(strip_private(rec), config=workload_config("espnet_b64_fp32", 64))
        rec = measure_crops(ctx, args, "espnet_b256_ens5", "f16tc", 256, max(3, args.steps // 4), 3, want_profile=False)
        free_rec(rec)
        extra["espnet_b256_ens5_f16tc"] = dict(strip_private(rec), config=workload_config("espnet_b256_ens5", 256))
        extra["batch1_latency"] = measure_batch1(ctx, args)
        torch.cuda.empty_cache()
        extra["sweep"] = measure_sweep(ctx, args)
        torch.cuda.empty_cache()
        w = measure_wsi(ctx, args, 1, 1)
        if w is not None:
            extra["wsi_%s" % args.slide] = w
    if ctx.rank != 0:
        ctx.close()
        return
    cpu_sample = 32 if wl != "espnet_b256_ens5" else 4
    cpu_val, cpu_dt = time_cpu(wl, cpu_sample, 4, 1) if ctx.world == 1 and not args.no_cpu else (None, None)
    value = head["value"]
    line = {
        "metric": "ESPNet 512x512 crops/s", "value": value, "unit": "crops/s", "n_gpus": ctx.world, "steps": args.steps,
        "warmup": warm, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE[args.mode], "data": "synthetic",
        "config": dict(workload_config(wl, B), mode=args.mode),
        "mpx_per_s": value * 512 * 512 / 1e6,
        "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"],
        "roofline": head.get("roofline"),
        "block_roofline_l3": head.get("block_roofline_l3"),
        "rooflines": head.get("rooflines"),
        "kernels": head.get("kernels"),
        "clocks": clocks,
        "host": ctx.host,
        "tflops_effective": value * (FLOP_PER_CROP_ENC if wl == "espnet_c_b64_fp32" else FLOP_PER_CROP_FULL) / 1e12,
    }
    for k in ("other_mode", "per_fold_mask_agreement_with_fp32_mode", "ensemble_mask_agreement_with_fp32_mode"):
        if k in head:
            line[k] = head[k]
    if cpu_val is not None:
        line["cpu_baseline"] = {"value": cpu_val, "unit": "crops/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d of the same synthetic 512x512 crops x 4 timed passes (%.1f s each), oracle port on torch %s CPU"
                                          % (cpu_sample, cpu_dt, torch.__version__)}
    if extra:
        line["configs"] = extra
    emit(line)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["espnet_c_b64_fp32", "espnet_b64_fp32", "espnet_b256_ens5", "wsi"],
                    help="default: espnet_c_b64_fp32 as the headline plus one nested record per remaining BASELINE config")
    ap.add_argument("--quick", action="store_true", help="headline workload only (no nested records)")
    ap.add_argument("--slide", default="40000x30000", help="wsi: synthetic slide WxH in level-0 pixels (BASELINE configs[3])")
    ap.add_argument("--overlap", type=float, default=0.1, help="wsi: tile overlap ratio (detect_glomus_test.py default)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="wsi at N > 1: how the band masks reach rank 0")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--depth", type=int, default=3, help="device slots of the host pipeline used for the e2e metric")
    ap.add_argument("--mode", default="fp32", choices=["fp32", "f16tc"],
                    help="fp32: fp32-equivalent path (1e-3 logit bar); f16tc: tcgen05 fp16-operand path (0.999 mask-agreement bar)")
    ap.add_argument("--fp32-impl", default="auto", choices=["auto", "cuda", "tc3"],
                    help="fp32 mode only: cuda = CUDA-core FMA kernels, tc3 = tcgen05 with 3-term fp16 operand splits (fp32-equivalent)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--single-mode", action="store_true", help="skip the second leg that times the other compute mode")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
