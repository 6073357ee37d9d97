#!/usr/bin/env python
"""bench.py -- OS-CNN + feature-level style-transfer training throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

Workload (N = 1): BASELINE.json configs[1] = "cfg2": target and source batches of B=128 series, C=9, L=128,
6 classes; one step = 2 extractors + DimensionUnification + AdaIN + Gram loss + 2 classifiers + CE, backward,
RMSprop.  For N > 1 every rank runs the same per-GPU batch on its own shard (weak scaling) and the
flat gradient bucket is all-reduced over NCCL.  `value` = series/s of the whole job with inputs resident in HBM,
`e2e` = the same with host->device copies of the step's inputs and a device->host read of the loss inside the
timed region.  One JSON line on stdout (rank 0); besides the contract's keys it carries `roofline` (in-step and isolated),
`roofline_hbm` / `roofline_gram`, `cpu_baseline`, `gpu_eager_baseline`, the per-kernel table and `step_ms_spread`.
`--scaling strong` shards the configuration's batch over the ranks instead (not the driver's contract).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "OS-CNN+style-transfer train samples/sec"
UNIT = "samples/s"
CFG = dict(name="cfg2", B=128, C=9, L=128, K=6)          # per domain, per GPU
WORKLOAD = "cfg2: OS-CNN + AdaIN/Gram style transfer, B=128 per domain per GPU, C=9, L=128, 6 classes"
# secondary workload (--workload cfg4, BASELINE configs[3]; not the headline line): long-series OS-CNN forward + backward
CFG4 = dict(name="cfg4", B=256, C=3, L=1024, K=4)
WORKLOAD4 = "cfg4: long-series OS-CNN extractor + classifier, forward + backward + RMSprop, B=256 per GPU, C=3, L=1024, primes up to 89"
# secondary workload (--workload cfg3, BASELINE configs[2]): three source domains + one target, style transfer + C-DAN
CFG3 = dict(name="cfg3", B=128, target=(9, 128, 6), sources=[(1, 128, 5), (3, 256, 4), (9, 128, 6)])
WORKLOAD3 = ("cfg3: 3 source domains (C,L)=(1,128),(3,256),(9,128) + 1 target (9,128), per-pair module sets, AdaIN/Gram style "
             "transfer + eval-BN classifier on the generated features + C-DAN loss, B=128 per domain per GPU")
STYLE_WEIGHT = 1.0


def step_config(workload: str, series_per_gpu: int):
    """`config` of the JSON line: workload keys only, identical in both arms (the driver compares them)."""
    return dict(workload=workload, series_per_step_per_gpu=series_per_gpu,
                l2="flushed between timed steps (256 MiB write, outside the per-step events)")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed region -- in process through NVML (`nvidia_ml_py`).  Round 1
    spawned `nvidia-smi` every 50 ms; with eight ranks doing that concurrently single steps of the timed region took 26 ms
    instead of 1.3 (`profiles/r2_dp_exchange.md`): a subprocess per sample is itself a disturbance.  `nvidia-smi` remains the
    fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.nvml, self.handle, self.source = None, None, "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(v) for v in vis.split(",")] if vis and all(v.strip().isdigit() for v in vis.split(",")) else None
            phys = ids[index] if ids and index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        return [str(sm), str(self.max_sm)] + ["Active" if mask & b else "Not Active" for b in bits]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self.nvml is not None:
                    try:
                        self.rows.append(self._sample_nvml())
                    except Exception:
                        self.nvml, self.source = None, "nvidia-smi"       # this NVML cannot answer: fall back for the rest
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.02 if self.nvml is not None else 0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=reasons,
                    samples=len(self.rows), source=self.source)


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the step on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_step_rate(steps: int, warmup: int):
    import torch
    from oracle import os_cnn as O
    from oracle import step as OS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ms = OS.ModelSet(CFG["C"], CFG["L"], CFG["K"], CFG["C"], CFG["L"], CFG["K"], seed=0)
    xt, yt = O.synthetic_batch(CFG["B"], CFG["C"], CFG["L"], CFG["K"], 0)
    xs, ys = O.synthetic_batch(CFG["B"], CFG["C"], CFG["L"], CFG["K"], 1)
    for _ in range(warmup):
        OS.train_step(ms, xt, yt, xs, ys, STYLE_WEIGHT)
    t0 = time.perf_counter()
    for _ in range(steps):
        OS.train_step(ms, xt, yt, xs, ys, STYLE_WEIGHT)
    dt = (time.perf_counter() - t0) / steps
    return 2 * CFG["B"] / dt, dt, cores, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # one step = one full cfg2 step of the oracle port (about 0.25 s on 16 cores): --steps / --warmup are honoured as given
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, dt, cores, threads = cpu_step_rate(steps, warmup)
    sample = f"{steps} full cfg2 steps (B=128 per domain) of the oracle port on {threads} torch CPU threads"
    line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=steps,
                warmup=warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic",
                config=step_config(WORKLOAD, 2 * CFG["B"]),
                run=dict(engine="cpu oracle port (torch/oneDNN fp32)", parallelism="rank 0 only"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)



def measure_tf32_peak(torch, n=8192, reps=6):
    """cuBLAS TF32 matmul throughput on this box (burst, best of `reps`), measured the way MEASURED_PEAKS.json measures
    the bf16 peak -- the denominator of the Gram kernels' tensor roofline (kind::tf32)."""
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(2):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def time_cold(torch, fn, flush, reps=8):
    """mean device ms of fn() with the L2 flushed before every launch (events on the launching stream)."""
    for _ in range(3):
        fn()
    evs = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def style_rooflines(torch, ops, L, peaks, tf32_peak, flush):
    """HBM roofline of the statistics / AdaIN kernels (north_star (b)) and tensor roofline of the Gram kernels (c), at the
    cfg2 shape and at the sweep shape [1024, 144, 1024]; every launch timed alone, L2 flushed before it (cold operands)."""
    hbm, gram = [], []
    for (B, C, Ln) in ((128, 144, 128), (1024, 144, 1024)):
        n = B * C * Ln
        c = torch.randn(B, C, Ln, device="cuda")
        st = torch.randn(B, C, Ln, device="cuda")
        dy = torch.randn(B, C, Ln, device="cuda")
        _, stats = ops.adain_fwd(c, st, 1e-5)
        for name, fn, nbytes in (("rowstats_welford", lambda: ops.rowstats(c), 4.0 * n),
                                 ("adain_fwd", lambda: ops.adain_fwd(c, st, 1e-5), 12.0 * n),
                                 ("adain_bwd", lambda: ops.adain_bwd(dy, c, st, stats), 20.0 * n)):
            ms = time_cold(torch, fn, flush)
            gbs = nbytes / (ms * 1e-3) / 1e9
            hbm.append(dict(kernel=name, shape=[B, C, Ln], bound="hbm", us=round(ms * 1e3, 2), achieved=round(gbs, 1),
                            peak=peaks["hbm"], unit="GB/s", frac=round(gbs / peaks["hbm"], 3), algorithmic_bytes=nbytes))
        eng = L.ENGINE_TCGEN05
        _, D = ops.gram_loss_fwd(eng, c, st)
        one = torch.ones((), device="cuda")
        flops = 2.0 * 2 * B * C * C * Ln
        for name, fn in (("gram_loss_fwd", lambda: ops.gram_loss_fwd(eng, c, st)),
                         ("gram_loss_bwd", lambda: ops.gram_loss_bwd(eng, D, c, st, one))):
            ms = time_cold(torch, fn, flush)
            tf = flops / (ms * 1e-3) / 1e12
            gram.append(dict(kernel=name, shape=[B, C, Ln], bound="tensor", us=round(ms * 1e3, 2), achieved=round(tf, 1),
                             peak=round(tf32_peak, 1), unit="TFLOP/s", frac=round(tf / tf32_peak, 3), algorithmic_flops=flops,
                             peak_source="cuBLAS TF32 8192^3 measured in this run (kind::tf32; the forward issues 3 MMAs per "
                                         "algorithmic one: hi*hi + hi*lo + lo*hi)"))
        del c, st, dy, stats, D
    return hbm, gram


def gpu_eager_baseline(torch, steps=10, warmup=3):
    """The GPU bar (SURVEY 8d / BASELINE.md section 4): the oracle port of the cfg2 step -- the reference's own operator
    sequence in plain PyTorch -- run eagerly on cuda:0 (cuDNN convolutions with torch's default TF32, cuBLAS, ATen
    element-wise kernels), timed with CUDA events after this repo's timed regions.  A reported baseline, like cpu_baseline."""
    from oracle import os_cnn as OO
    from oracle import step as OS
    ms = OS.ModelSet(CFG["C"], CFG["L"], CFG["K"], CFG["C"], CFG["L"], CFG["K"], seed=0)
    for name in ("fe_t", "cl_t", "fe_s", "du", "cl_s"):
        sd = getattr(ms, name)
        for k in list(sd.keys()):
            sd[k] = sd[k].cuda()
    xt, yt = [t.cuda() for t in OO.synthetic_batch(CFG["B"], CFG["C"], CFG["L"], CFG["K"], 0)]
    xs, ys = [t.cuda() for t in OO.synthetic_batch(CFG["B"], CFG["C"], CFG["L"], CFG["K"], 1)]

    def one():
        ms.set_requires_grad()
        out = OS.step_forward(ms, xt, yt, xs, ys, STYLE_WEIGHT, training=True)
        out["loss"].backward()
        OS.rmsprop_update(ms)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / steps
    return dict(value=2 * CFG["B"] / dt, unit=UNIT, ms_per_step=dt * 1e3, steps=steps,
                what="oracle port of the cfg2 step in eager PyTorch on cuda:0: cuDNN convolutions (TF32, torch default), "
                     "cuBLAS fp32, ATen element-wise kernels, one optimizer loop over the tensors; host launch latency included "
                     "(that is how the reference runs)")


# ----------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------
class KernelProfile:
    """Times every C-ABI call family with CUDA events on the launching stream (instrumented pass, run after
    the timed region; never part of a reported step time)."""
    LAUNCHES = dict(ncl_to_c8=1, c8_to_ncl=1, pack_weights=1, pack_weights_pair=1, pack_weights_multi=1, rmsprop_step=1,
                    osconv=1, oswgrad=2, head_ce_fwd=1, head_ce_bwd=1, weighted_scalar_sum=1, bn_stats=2, bn_eval_coeffs=1, bn_apply=1, bn_bwd_reduce=2, bn_bwd_apply=1,
                    bn_apply_fused=1, bn_bwd_top=1, bn_bwd_top_pooled=1, bn_bwd_apply_fused=1, adain_fwd=1, adain_bwd=1, gram_loss_fwd=2,
                    gram_loss_bwd=1, rowstats=1, multi_l2norm=2, class_precision=1, entropy_vote=1, cdan_fuse_fwd=1, cdan_fuse_bwd=1, cdan_distance_fwd=1, cdan_distance_bwd=1)

    def __init__(self, ops, torch):
        self.ops, self.torch, self.records, self.orig, self.count = ops, torch, [], {}, 0
        self.timing = False
        self.conv_calls = None          # while a list: (direction, geometry, B, L, fused epilogue?) of every osconv call

    def _meta(self, name, args):
        L = self.ops.L
        if name == "osconv":
            engine, direction, g, x8 = args[0], args[1], args[2], args[3]
            B, _, Ln, _ = x8.shape
            return dict(flops=2.0 * B * Ln * g.live_macs_per_position(), tc=engine == L.ENGINE_TCGEN05)
        if name == "oswgrad":
            engine, g, dy8, x8 = args[0], args[1], args[2], args[3]
            B, _, Ln, _ = x8.shape
            return dict(flops=2.0 * B * Ln * g.live_macs_per_position(), tc=engine == L.ENGINE_TCGEN05)
        if name == "adain_fwd":
            return dict(bytes=3.0 * args[0].numel() * 4)
        if name == "adain_bwd":
            return dict(bytes=5.0 * args[0].numel() * 4)
        if name == "gram_loss_fwd":
            B, C, Ln = args[1].shape
            return dict(flops=2.0 * 2 * B * C * C * Ln, tc=args[0] == L.ENGINE_TCGEN05)
        if name == "gram_loss_bwd":
            B, C, Ln = args[2].shape
            return dict(flops=2.0 * 2 * B * C * C * Ln, tc=args[0] == L.ENGINE_TCGEN05)
        if name in ("bn_apply", "bn_stats", "bn_bwd_reduce", "bn_bwd_apply"):
            y8 = args[0]
            n = y8.numel() * 4.0
            mult = dict(bn_stats=1.0, bn_apply=1.75, bn_bwd_reduce=2.0, bn_bwd_apply=2.5)[name]
            return dict(bytes=n * mult)
        if name == "bn_apply_fused":          # read y (fp32) [+ second branch], write bf16 c8 or fp32 NCL
            n = args[0].y8.numel() * 4.0
            return dict(bytes=n * ((2.0 if args[1] is not None else 1.0) + (1.0 if args[4] == L.OUT_NCL_F32 else 0.5)))
        if name == "bn_bwd_top":               # read dout, y [+ y2], write d
            n = args[1].y8.numel() * 4.0
            return dict(bytes=n * (3.0 + (1.0 if args[2] is not None else 0.0)))
        if name == "bn_bwd_apply_fused":       # read d, y, write dy (bf16)
            n = args[0].numel() * 4.0
            return dict(bytes=n * 2.5)
        return {}

    def install(self):
        for name in self.LAUNCHES:
            fn = getattr(self.ops, name)
            self.orig[name] = fn

            def wrapped(*a, __fn=fn, __name=name, **k):
                self.count += self.LAUNCHES[__name]
                if __name == "pack_weights" and len(a) > 4 and a[4]:
                    self.count += 1
                if __name == "osconv" and self.conv_calls is not None:
                    self.conv_calls.append((a[1], a[2], a[3].shape[0], a[3].shape[2],
                                            k.get("stat_partial") is not None, k.get("red_partial") is not None))
                if not self.timing:
                    return __fn(*a, **k)
                e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
                e0.record()
                r = __fn(*a, **k)
                e1.record()
                self.records.append((__name, e0, e1, self._meta(__name, a)))
                return r
            setattr(self.ops, name, wrapped)

    def uninstall(self):
        for name, fn in self.orig.items():
            setattr(self.ops, name, fn)

    def table(self, steps):
        agg = {}
        for name, e0, e1, meta in self.records:
            key = name + ("[tc]" if meta.get("tc") else "")
            d = agg.setdefault(key, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
            d["ms"] += e0.elapsed_time(e1)
            d["n"] += 1
            d["flops"] += meta.get("flops", 0.0)
            d["bytes"] += meta.get("bytes", 0.0)
        for d in agg.values():
            d["ms_per_step"] = d["ms"] / steps
            d["launches_per_step"] = d["n"] / steps
            d["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["flops"] else None
            d["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["bytes"] else None
        return agg


def conv_traffic_from_profile(launches_per_step):
    """dram__bytes_read.sum + dram__bytes_write.sum of the step's conv launches (osconv2_kernel) from the committed ncu capture of
    the eager cfg2 step (profiles/r2_step_dram_traffic.json, made by tools/ncu_traffic.py from profiles/r2_traffic_step.csv).
    Bytes per step, like `achieved`."""
    path = os.path.join(ROOT, "profiles", "r2_step_dram_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed"
    with open(path) as f:
        kernels = json.load(f)["kernels"]
    for name, d in kernels.items():
        if "osconv2_kernel<0, 0, 0>" in name and abs(d["launches_per_step"] - launches_per_step) < 0.5:
            return (d["dram_read_bytes_per_step"] + d["dram_write_bytes_per_step"],
                    "bytes per step over the %d conv launches, ncu cold-cache capture of the eager cfg2 step "
                    "(profiles/r2_step_dram_traffic.json): reads only -- at this size every write stays in the 126 MB L2"
                    % launches_per_step)
    return None, "the committed ncu capture is of the cfg2 step (24 conv launches); this workload differs"


def conv_roofline(torch, ops, L, conv_calls, peaks, reps=20):
    """Roofline of the dominant kernel (osconv_tc_kernel, forward + dgrad): every distinct (bank, direction, epilogue)
    launch of the step is replayed in isolation -- `reps` back-to-back launches between two CUDA events on the launching
    stream (captured in a CUDA graph, so no host latency sits between them), operands L2-warm as they are inside the
    step -- and weighted by its call count.
    achieved = live-tap FLOPs of all conv launches of one step / their summed device time."""
    dev = torch.device("cuda", torch.cuda.current_device())
    groups = {}
    for (direction, g, B, Ln, stat, red) in conv_calls:
        key = (direction, id(g), B, Ln, stat, red)
        groups.setdefault(key, [g, 0])[1] += 1
    tot_t = tot_f = 0.0
    detail = []
    for (direction, _, B, Ln, stat, red), (g, count) in groups.items():
        fwd = direction == L.DIR_FWD
        cin_side, cout_side = (g.cin, g.cout) if fwd else (g.cout, g.cin)
        x8 = ops.ncl_to_c8(torch.randn(B, cin_side, Ln, device=dev), L.TSC_BF16)
        W = torch.randn(g.cout, g.cin, g.kmax, device=dev) * 0.05
        wp = ops.pack_weights(g, W, direction, L.TSC_BF16, False)
        bias = torch.zeros(g.cout, device=dev) if fwd else None
        ncta = ops.n_conv_ctas(B, Ln)
        kw = {}
        if stat:
            kw["stat_partial"] = torch.empty(ncta, ops.pad16(cout_side), 2, device=dev)
        if red:
            cp = ops.pad16(cout_side)
            yb = torch.randn(B, cp // 8, Ln, 8, device=dev)
            one = torch.ones(cp, device=dev)
            kw["mask"] = (yb, one, one * 0.1, one * 0.0, one)
            kw["red_partial"] = torch.empty(ncta, cp, 2, device=dev)
        for _ in range(3):
            ops.osconv(L.ENGINE_TCGEN05, direction, g, x8, wp, bias, **kw)
        torch.cuda.synchronize()
        # `reps` launches captured in a CUDA graph: the events then bracket device time only (no host launch latency)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.osconv(L.ENGINE_TCGEN05, direction, g, x8, wp, bias, **kw)
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(graph):
            for _ in range(reps):
                ops.osconv(L.ENGINE_TCGEN05, direction, g, x8, wp, bias, **kw)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        f = 2.0 * B * Ln * g.live_macs_per_position()
        tot_t += t * count
        tot_f += f * count
        detail.append(dict(direction="fwd" if fwd else "dgrad", cin=g.cin, cout=g.cout, kmax=g.kmax, calls=count,
                           us=round(t * 1e6, 2), tflops=round(f / t / 1e12, 1)))
    achieved = tot_f / tot_t / 1e12 if tot_t else 0.0
    traffic, traffic_note = conv_traffic_from_profile(sum(c for _, c in groups.values()))
    return dict(kernel="osconv2_kernel (forward + dgrad launches of one step)", bound="tensor", achieved=achieved,
                peak=peaks["bf16"], unit="TFLOP/s", frac=achieved / peaks["bf16"], traffic=traffic, traffic_note=traffic_note,
                peak_source=peaks["source"] + " bf16 burst (kernel timed alone)",
                algorithmic_flops_per_step=tot_f, device_ms_per_step=tot_t * 1e3, launches=detail)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import feature_level_style_transfer_for_tsc_b200 as T
    from feature_level_style_transfer_for_tsc_b200 import ops
    from feature_level_style_transfer_for_tsc_b200.train_step import StyleTransferModelSet, Trainer
    from feature_level_style_transfer_for_tsc_b200 import data as O      # synthetic input generator (SURVEY 8d definition)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the tsc_b200 kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"             # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    T._lib.load()
    T.set_engine(args.engine)

    torch.manual_seed(0)

    def per_gpu(b_total: int) -> int:
        """weak scaling (default): every rank runs the configuration's batch; strong scaling (--scaling strong, BASELINE
        configs[3] "batch-sharded at 2/4/8 GPUs"): the configuration's batch is sharded over the ranks."""
        if args.scaling == "weak":
            return b_total
        if b_total % world != 0:
            raise SystemExit(f"--scaling strong: the batch of {b_total} series does not divide over {world} ranks")
        return b_total // world

    cfg4 = args.workload == "cfg4"
    cfg3 = args.workload == "cfg3"
    if cfg3:
        from feature_level_style_transfer_for_tsc_b200.train_step import MultiSourceModelSet
        model = MultiSourceModelSet(CFG3["target"], CFG3["sources"]).to(dev)
        B = per_gpu(CFG3["B"])
        n_dom = 1 + len(CFG3["sources"])
        batches = [O.synthetic_batch(B, *CFG3["target"], n_dom * rank)]
        batches += [O.synthetic_batch(B, C, Ln, K, n_dom * rank + 1 + i) for i, (C, Ln, K) in enumerate(CFG3["sources"])]
        host = [t.pin_memory() for xb in batches for t in xb]
        series_per_gpu = n_dom * B
    elif cfg4:
        from feature_level_style_transfer_for_tsc_b200.train_step import SingleDomainModelSet
        model = SingleDomainModelSet(CFG4["C"], CFG4["L"], CFG4["K"]).to(dev)
        B = per_gpu(CFG4["B"])
        x_h, y_h = O.synthetic_batch(B, CFG4["C"], CFG4["L"], CFG4["K"], rank)
        host = [t.pin_memory() for t in (x_h, y_h)]
        series_per_gpu = B
    else:
        model = StyleTransferModelSet(CFG["C"], CFG["L"], CFG["K"], CFG["C"], CFG["L"], CFG["K"]).to(dev)
        B = per_gpu(CFG["B"])
        xt_h, yt_h = O.synthetic_batch(B, CFG["C"], CFG["L"], CFG["K"], 2 * rank)
        xs_h, ys_h = O.synthetic_batch(B, CFG["C"], CFG["L"], CFG["K"], 2 * rank + 1)
        host = [t.pin_memory() for t in (xt_h, yt_h, xs_h, ys_h)]
        series_per_gpu = 2 * B
    trainer = Trainer(model, STYLE_WEIGHT, use_graph=not args.no_graph)
    trainer.broadcast_parameters(0)
    dev_in = [t.to(dev) for t in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)      # 256 MiB > 126 MB L2
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    prof = KernelProfile(ops, torch)
    prof.install()
    # one eager step to count this repo's kernel launches per step (the same launches the CUDA graph replays)
    trainer.use_graph = False
    n0 = prof.count
    prof.conv_calls = []
    trainer.step(*dev_in)
    launches = prof.count - n0
    conv_calls, prof.conv_calls = prof.conv_calls, None
    trainer.use_graph = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        evs = []
        for _ in range(nsteps):
            flush.zero_()                                           # L2 flush, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if e2e:
                ins = [h.to(dev, non_blocking=True) for h in host]
                loss = trainer.step(*ins)
                loss_host.copy_(loss, non_blocking=True)
            else:
                trainer.step(*dev_in)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        spread.append((ts[0], ts[len(ts) // 2], ts[-1]))
        return sum(ts) * 1e-3

    spread = []                                  # (min, median, max) step ms of every timed region of this rank
    for _ in range(max(args.warmup, 3)):
        trainer.step(*dev_in)
    if world > 1:
        # a captured NCCL collective keeps setting itself up over its first replays (measured on 8 GPUs: the first ~100
        # replays of the step graph averaged 2.1 ms against 1.3 ms afterwards): these extra untimed replays come on top of
        # the --warmup steps, the K timed steps are unchanged
        for _ in range(30):
            trainer.step(*dev_in)
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_dev = timed(args.steps, False)
    barrier()
    t_e2e = timed(args.steps, True)
    barrier()
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    series = series_per_gpu * world * args.steps

    # instrumented pass (per-kernel-family device time; not part of any reported step time)
    prof.timing = True
    trainer.use_graph = False
    # one stream for this pass: an event pair around a launch must not contain waits on the other branch's stream
    for m in model.modules():
        if hasattr(m, "two_streams"):
            m.two_streams = False
        if hasattr(m, "multi_stream"):
            m.multi_stream = False
    psteps = 3
    # keep the GPU busy with memsets while the host enqueues the whole eager step, so that every event pair brackets
    # device time only -- without this the host's launch latency sits between the events.  The backlog is sized from
    # the measured host time of one eager step and the measured duration of one memset (x1.5).
    prof.timing = False
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trainer.step(*dev_in)
    host_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        flush.zero_()
    e1.record()
    torch.cuda.synchronize()
    memset_s = e0.elapsed_time(e1) * 1e-3 / 20
    backlog = int(1.5 * host_s / memset_s) + 50
    prof.timing = True
    for _ in range(psteps):
        for _ in range(backlog):
            flush.zero_()
        trainer.step(*dev_in)
    torch.cuda.synchronize()
    table = prof.table(psteps)
    prof.uninstall()

    def shutdown():
        """Leave the process group without ever hanging the launcher: the step graph (NCCL inside) is released first, and a
        watchdog ends the process if the teardown still blocks."""
        if world > 1:
            sys.stdout.flush()
            t = threading.Timer(30.0, lambda: os._exit(0))
            t.daemon = True
            t.start()
            trainer.release_graph()
            dist.destroy_process_group()
            t.cancel()

    if rank != 0:
        shutdown()
        return
    peaks = load_peaks()
    kern = {k: dict(ms_per_step=round(v["ms_per_step"], 4), launches=v["launches_per_step"],
                    tflops=(round(v["tflops"], 2) if v["tflops"] else None),
                    gbs=(round(v["gbs"], 1) if v["gbs"] else None)) for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])}
    extra = {}
    if args.engine == "tcgen05":
        iso = conv_roofline(torch, ops, T._lib, conv_calls, peaks)
        # headline = the kernel INSIDE the step: algorithmic FLOPs of the step's conv launches / their summed event time in
        # the instrumented single-stream pass, against the sustained bf16 peak (a kernel timed inside a long step)
        d = table.get("osconv[tc]")
        ach = d["tflops"] if d and d["tflops"] else 0.0
        roof = dict(kernel=iso["kernel"], bound="tensor", achieved=ach, peak=peaks["bf16_sustained"], unit="TFLOP/s",
                    frac=ach / peaks["bf16_sustained"], traffic=iso["traffic"], traffic_note=iso["traffic_note"],
                    peak_source=peaks["source"] + " bf16 sustained (kernel timed inside the step)",
                    timing="CUDA events around every conv launch of 3 instrumented eager steps on one stream (device time, GPU "
                           "kept busy ahead of the host); FLOPs = live-tap 2*B*L*Cin*sum(out_g*k_g) per launch",
                    algorithmic_flops_per_step=iso["algorithmic_flops_per_step"],
                    device_ms_per_step=(d["ms_per_step"] if d else None), launches_per_step=(d["launches_per_step"] if d else None),
                    isolated=dict(achieved=iso["achieved"], peak=iso["peak"], frac=iso["frac"], peak_source=iso["peak_source"],
                                  device_ms_per_step=iso["device_ms_per_step"], launches=iso["launches"],
                                  note="each distinct launch replayed 20x back to back in a CUDA graph, operands L2-warm"))
        if world == 1 and not args.no_extra:
            tf32_peak = measure_tf32_peak(torch)
            hbm, gram = style_rooflines(torch, ops, T._lib, peaks, tf32_peak, flush)
            extra = dict(roofline_hbm=hbm, roofline_gram=gram, tf32_peak_tflops=round(tf32_peak, 1))
    else:
        d = table.get("osconv", dict(tflops=0.0))
        roof = dict(kernel="osconv_simt_kernel (fp32 CUDA cores; not the product engine)", bound="tensor",
                    achieved=d["tflops"] or 0.0, peak=peaks["bf16"], unit="TFLOP/s", frac=(d["tflops"] or 0.0) / peaks["bf16"],
                    traffic=None, peak_source=peaks["source"] + " bf16 burst")
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not cfg4 and not cfg3:
        v, dt, cores, threads = cpu_step_rate(3, 1)
        cpu = dict(value=v, unit=UNIT, cores=threads, kind="port",
                   sample=f"3 full cfg2 steps (B=128 per domain) of the oracle port, {dt * 1e3:.0f} ms/step, {cores} host cores")
    line = dict(metric=METRIC, value=series / t_dev, unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=t_dev / args.steps * 1e3, higher_is_better=True, scaling=args.scaling,
                vs_baseline=None, dtype="bf16" if args.engine == "tcgen05" else "f32", data="synthetic",
                config=step_config((WORKLOAD3 if cfg3 else WORKLOAD4 if cfg4 else WORKLOAD)
                                   + (f" [strong scaling: that batch sharded over {world} ranks]" if args.scaling == "strong" else ""),
                                   series_per_gpu),
                run=dict(engine=args.engine, parallelism=f"dp{world}", cuda_graph=not args.no_graph),
                e2e=dict(value=series / t_e2e, unit=UNIT, h2d_bytes_per_step=h2d_bytes * world, d2h_bytes_per_step=4 * world,
                         ms_per_step=t_e2e / args.steps * 1e3),
                gpu_launches=int(launches), clocks=clocks.summary(), roofline=roof, kernels=kern,
                step_ms_spread=dict(device=dict(zip(("min", "median", "max"), [round(v, 4) for v in spread[0]])),
                                    e2e=dict(zip(("min", "median", "max"), [round(v, 4) for v in spread[1]])),
                                    note="rank 0, per-step CUDA-event times of the two timed regions"))
    if cpu:
        line["cpu_baseline"] = cpu
    line.update(extra)
    if world == 1 and not args.no_extra and not cfg4 and not cfg3:
        line["gpu_eager_baseline"] = gpu_eager_baseline(torch)
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the statistics / AdaIN / Gram rooflines, the TF32 peak and the eager-PyTorch GPU baseline")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4"],
                    help="cfg2 = the headline step (default); cfg3 = multi-source transfer with the C-DAN loss; "
                         "cfg4 = long-series OS-CNN forward + backward (both secondary)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): fixed per-GPU batch; strong: the configuration's batch is "
                         "sharded over the ranks (BASELINE configs[3]: cfg4 B=256 at 2/4/8 GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
