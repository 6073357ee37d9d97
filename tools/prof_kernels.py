#!/usr/bin/env python
"""Runs the hot kernels of the cfg2 step in isolation (for `ncu --set full -k regex:...` captures and for
CUDA-event timing of a single kernel):  python tools/prof_kernels.py [--layer cfg2_l1] [--B 128] [--iters 5]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_level_style_transfer_for_tsc_b200 as T                     # noqa: E402
from feature_level_style_transfer_for_tsc_b200 import ops                 # noqa: E402
from feature_level_style_transfer_for_tsc_b200.train_step import trainer_layer_lists   # noqa: E402

L = T._lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=9)
    ap.add_argument("--L", type=int, default=128)
    ap.add_argument("--B", type=int, default=128)
    ap.add_argument("--layer", type=int, default=1, help="index into the extractor's layer list (or 3.. = classifier)")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--stages", action="store_true", help="print the per-stage timeline of CTA 0")
    a = ap.parse_args()
    ext, cls, cf = trainer_layer_lists(a.C, a.L)
    layers = ext + cls
    g = ops.bank_geometry(layers[a.layer])
    dev = "cuda"
    x = torch.randn(a.B, g.cin, a.L, device=dev)
    dy = torch.randn(a.B, g.cout, a.L, device=dev)
    W = torch.randn(g.cout, g.cin, g.kmax, device=dev) * 0.05
    bias = torch.randn(g.cout, device=dev)
    x8 = ops.ncl_to_c8(x, L.TSC_BF16)
    dy8 = ops.ncl_to_c8(dy, L.TSC_BF16)
    wf, wd = ops.pack_weights_pair(g, W, L.TSC_BF16, True, True)
    flops = 2.0 * a.B * a.L * g.live_macs_per_position()
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    res = {}
    tl = torch.zeros(8 + 48 * 8, device=dev, dtype=torch.int64)
    names = ["entry", "setup", "x tile", "first W", "MMAs issued", "acc ready", "epilogue", "exit"]
    for name, fn in (("fwd", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, g, x8, wf, bias)),
                     ("dgrad", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_DGRAD, g, dy8, wd, None)),
                     ("wgrad", lambda: ops.oswgrad(L.ENGINE_TCGEN05, g, dy8, x8))):
        fn(); torch.cuda.synchronize()
        L.load().tsc_debug_set_timeline(tl.data_ptr())
        fn(); torch.cuda.synchronize()
        L.load().tsc_debug_set_timeline(None)
        t = tl.cpu().tolist()
        print(f"timeline[{name}] CTA0 cycles since entry: " + ", ".join(f"{n}={v - t[0]}" for n, v in zip(names, t)))
        if a.stages and name != "wgrad":
            print("  stage: issuer[wait-begin, both halves landed, -, issued+committed]  producer[wait-begin, slot free]")
            for i in range(48):
                r = t[8 + i * 8: 16 + i * 8]
                if r[1] == 0:
                    break
                print(f"   {i:3d}: " + " ".join(f"{v - t[0]:7d}" for v in r[0:4]) + "   | " + " ".join(f"{v - t[0]:7d}" for v in r[4:6]))
        if name == "wgrad":
            it = [v - t[0] for v in t[8:32] if v]
            print("  wgrad tile-ready times (cycles since entry): " + " ".join(str(v) for v in it))
        tl.zero_()
    for name, fn in (("fwd", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, g, x8, wf, bias)),
                     ("dgrad", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_DGRAD, g, dy8, wd, None)),
                     ("wgrad", lambda: ops.oswgrad(L.ENGINE_TCGEN05, g, dy8, x8))):
        evs = []
        torch.cuda.synchronize()
        for i in range(a.iters + 3):
            # the 256 MB flush keeps the GPU busy while the host enqueues the next launch, so the event pair
            # brackets device time only (no host launch latency inside it); no sync inside the loop
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = [e0.elapsed_time(e1) * 1e3 for e0, e1 in evs[3:]]
        us = sorted(ts)[len(ts) // 2]
        res[name] = us
        print(f"{name:6s} Cin={g.cin} Cout={g.cout} Kmax={g.kmax} B={a.B} L={a.L}: {us:8.1f} us  "
              f"{flops / us / 1e6:8.1f} TFLOP/s (live-tap FLOPs {flops / 1e9:.2f} G)")
    assert ops.read_watchdog() == 0


if __name__ == "__main__":
    main()
