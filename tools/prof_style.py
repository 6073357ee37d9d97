#!/usr/bin/env python
"""The statistics / AdaIN / Gram kernels alone at one shape, for `ncu --set full -k regex:...` captures and CUDA-event
timing:  python tools/prof_style.py [--B 1024 --C 144 --L 1024]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_level_style_transfer_for_tsc_b200 as T                     # noqa: E402
from feature_level_style_transfer_for_tsc_b200 import ops                 # noqa: E402

L = T._lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--C", type=int, default=144)
    ap.add_argument("--L", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    x = torch.randn(a.B, a.C, a.L, device="cuda")
    s = torch.randn(a.B, a.C, a.L, device="cuda")
    dy = torch.randn(a.B, a.C, a.L, device="cuda")
    one = torch.ones((), device="cuda")
    nbytes = x.numel() * 4
    flops = 2.0 * 2 * a.B * a.C * a.C * a.L
    fns = [("rowstats_welford", lambda: ops.rowstats(x), nbytes, None),
           ("adain_fwd", lambda: ops.adain_fwd(x, s, 1e-5), 3 * nbytes, None)]
    out, stats = ops.adain_fwd(x, s, 1e-5)
    fns.append(("adain_bwd", lambda: ops.adain_bwd(dy, x, s, stats), 5 * nbytes, None))
    loss, D = ops.gram_loss_fwd(L.ENGINE_TCGEN05, x, s)
    fns.append(("gram_loss_fwd", lambda: ops.gram_loss_fwd(L.ENGINE_TCGEN05, x, s), None, flops))
    fns.append(("gram_loss_bwd", lambda: ops.gram_loss_bwd(L.ENGINE_TCGEN05, D, x, s, one), None, flops))
    for name, fn, by, fl in fns:
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        rate = f"{by / us / 1e3:8.1f} GB/s" if by else f"{fl / us / 1e6:8.1f} TFLOP/s"
        print(f"{name:18s} [{a.B},{a.C},{a.L}] {us:9.1f} us {rate}")


if __name__ == "__main__":
    main()
