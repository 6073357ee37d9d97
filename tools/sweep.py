#!/usr/bin/env python
"""cfg5 sweep: every hot kernel timed alone (CUDA events around a graph of back-to-back launches, inputs larger than
or comparable to L2 where the shape allows) over batch / length, reported against the measured peaks:
tensor kernels in TFLOP/s of live-tap (algorithmic) FLOPs vs the bf16 burst peak, streaming kernels in GB/s of
algorithmic bytes vs the HBM copy peak.   python tools/sweep.py [--quick] > gpurun_out/sweep.jsonl"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_level_style_transfer_for_tsc_b200 as T                     # noqa: E402
from feature_level_style_transfer_for_tsc_b200 import ops                 # noqa: E402
from feature_level_style_transfer_for_tsc_b200.train_step import trainer_layer_lists   # noqa: E402

L = T._lib


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    hbm, bf16, src = peaks()
    dev = "cuda"
    out = []

    def emit(**kw):
        out.append(kw)
        print(json.dumps(kw), flush=True)

    # ---- conv family on the widest bank of each configuration ----
    conv_shapes = [(9, 128, 128), (9, 128, 1024), (9, 128, 4096)] if a.quick else \
        [(9, 128, 128), (9, 128, 1024), (9, 128, 4096), (3, 1024, 256), (3, 1024, 1024), (9, 512, 1024), (3, 4096, 64)]
    for (C, Ln, B) in conv_shapes:
        ext, cls, cf = trainer_layer_lists(C, Ln)
        g = ops.bank_geometry(ext[1])
        x8 = ops.ncl_to_c8(torch.randn(B, g.cin, Ln, device=dev), L.TSC_BF16)
        dy8 = ops.ncl_to_c8(torch.randn(B, g.cout, Ln, device=dev), L.TSC_BF16)
        W = torch.randn(g.cout, g.cin, g.kmax, device=dev) * 0.05
        wf, wd = ops.pack_weights_pair(g, W, L.TSC_BF16, True, True)
        bias = torch.zeros(g.cout, device=dev)
        part = torch.empty(ops.n_conv_ctas(B, Ln), g.cout_p, 2, device=dev)
        flops = 2.0 * B * Ln * g.live_macs_per_position()
        reps = 5 if B * Ln >= 1 << 19 else 20
        gamma, beta = torch.ones(g.cout, device=dev), torch.zeros(g.cout, device=dev)
        rmean, rvar = torch.zeros(g.cout, device=dev), torch.ones(g.cout, device=dev)
        aff = ((gamma, beta, rmean, rvar, 1e-5), True, L.OUT_C8_BF16, None)
        for name, fn in (("osconv fwd (+BN partial statistics)", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, g, x8, wf, bias, stat_partial=part)),
                         ("osconv fwd inference (eval BN + ReLU epilogue, bf16 out)", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, g, x8, wf, bias, affine=aff)),
                         ("osconv dgrad", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_DGRAD, g, dy8, wd, None)),
                         ("oswgrad (+ordered reduce)", lambda: ops.oswgrad(L.ENGINE_TCGEN05, g, dy8, x8))):
            t = timed(fn, reps)
            emit(kernel=name, shape=f"B={B} L={Ln} {g.cin}->{g.cout} Kmax={g.kmax}", us=round(t * 1e6, 1), bound="tensor",
                 achieved=round(flops / t / 1e12, 1), unit="TFLOP/s", peak=bf16, frac=round(flops / t / 1e12 / bf16, 3), peak_source=src)
        del x8, dy8, part
    # ---- streaming kernels: AdaIN / row statistics / fused BN apply, and the Gram loss ----
    row_shapes = [(128, 144, 128), (1024, 144, 1024)] if a.quick else [(128, 144, 128), (1024, 144, 128), (1024, 144, 1024), (256, 50, 4096), (4096, 144, 128)]
    for (B, C, Ln) in row_shapes:
        c, s = torch.randn(B, C, Ln, device=dev), torch.randn(B, C, Ln, device=dev)
        dy = torch.randn(B, C, Ln, device=dev)
        n = c.numel() * 4.0
        reps = 5 if n > 2e8 else 20
        o, st = ops.adain_fwd(c, s, 1e-5)
        for name, fn, nbytes in (("rowstats_welford", lambda: ops.rowstats(c), n),
                                 ("adain_fwd", lambda: ops.adain_fwd(c, s, 1e-5), 3 * n),
                                 ("adain_bwd", lambda: ops.adain_bwd(dy, c, s, st), 5 * n)):
            t = timed(fn, reps)
            emit(kernel=name, shape=f"[{B},{C},{Ln}] fp32", us=round(t * 1e6, 1), bound="hbm", achieved=round(nbytes / t / 1e9, 1),
                 unit="GB/s", peak=hbm, frac=round(nbytes / t / 1e9 / hbm, 3), peak_source=src)
        if C <= 256:
            gf = 2.0 * 2 * B * C * C * Ln
            loss, D = ops.gram_loss_fwd(L.ENGINE_TCGEN05, c, s)
            one = torch.ones((), device=dev)
            for name, fn in (("gram_loss_fwd (3xTF32)", lambda: ops.gram_loss_fwd(L.ENGINE_TCGEN05, c, s)),
                             ("gram_loss_bwd (TF32)", lambda: ops.gram_loss_bwd(L.ENGINE_TCGEN05, D, c, s, one))):
                t = timed(fn, reps)
                emit(kernel=name, shape=f"[{B},{C},{Ln}] fp32", us=round(t * 1e6, 1), bound="tensor", achieved=round(gf / t / 1e12, 1),
                     unit="TFLOP/s (algorithmic)", peak=bf16 / 2, frac=round(gf / t / 1e12 / (bf16 / 2), 3), peak_source=src + " bf16 burst / 2 (TF32)")
        del c, s, dy, o, st
    assert ops.read_watchdog() == 0
    # markdown table
    print("\n| kernel | shape | us | achieved | peak | frac |\n|---|---|---:|---:|---:|---:|", file=sys.stderr)
    for r in out:
        print(f"| {r['kernel']} | {r['shape']} | {r['us']} | {r['achieved']} {r['unit']} | {r['peak']} | {r['frac']:.3f} |", file=sys.stderr)


if __name__ == "__main__":
    main()
