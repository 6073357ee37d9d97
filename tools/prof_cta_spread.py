#!/usr/bin/env python
"""Where does a conv launch's wall time go beyond the lifetime of one CTA?  Every CTA of the instrumented instantiation
records %globaltimer at entry / exit, its SM and its clock64 lifetime (TSC_CONV_DEBUG=16):
    TSC_CONV_DEBUG=16 python tools/prof_cta_spread.py [--B 128] [--layer 1]"""
import argparse
import os
import sys

os.environ.setdefault("TSC_CONV_DEBUG", "16")
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_level_style_transfer_for_tsc_b200 as T                      # noqa: E402
from feature_level_style_transfer_for_tsc_b200 import ops                  # noqa: E402
from feature_level_style_transfer_for_tsc_b200.train_step import trainer_layer_lists   # noqa: E402

L = T._lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=128)
    ap.add_argument("--layer", type=int, default=1)
    a = ap.parse_args()
    ext, cls, cf = trainer_layer_lists(9, 128)
    g = ops.bank_geometry((ext + cls)[a.layer])
    dev = "cuda"
    x8 = ops.ncl_to_c8(torch.randn(a.B, g.cin, 128, device=dev), L.TSC_BF16)
    dy8 = ops.ncl_to_c8(torch.randn(a.B, g.cout, 128, device=dev), L.TSC_BF16)
    W = torch.randn(g.cout, g.cin, g.kmax, device=dev) * 0.05
    wf, wd = ops.pack_weights_pair(g, W, L.TSC_BF16, True, True)
    bias = torch.zeros(g.cout, device=dev)
    n = ops.n_conv_ctas(a.B, 128)
    tl = torch.zeros(1024 + 4 * n, device=dev, dtype=torch.int64)
    for name, fn in (("fwd", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, g, x8, wf, bias)),
                     ("dgrad", lambda: ops.osconv(L.ENGINE_TCGEN05, L.DIR_DGRAD, g, dy8, wd, None))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        L.load().tsc_debug_set_timeline(tl.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        L.load().tsc_debug_set_timeline(None)
        r = tl[1024:].cpu().numpy().reshape(n, 4)
        t0 = r[:, 0].min()
        ent, ext_, life = (r[:, 0] - t0) / 1e3, (r[:, 1] - t0) / 1e3, (r[:, 1] - r[:, 0]) / 1e3
        q = lambda v: "min %.2f  p50 %.2f  p90 %.2f  max %.2f" % (v.min(), np.percentile(v, 50), np.percentile(v, 90), v.max())
        print(f"{name}: {n} CTAs on {len(set(r[:, 2].tolist()))} SMs; event time of the launch {e0.elapsed_time(e1) * 1e3:.1f} us (instrumented kernel)")
        print("   entry  [us after the first CTA]: " + q(ent))
        print("   exit   [us after the first CTA]: " + q(ext_))
        print("   CTA lifetime [us]:               " + q(life))
        print("   CTA lifetime [k cycles]:         " + q(r[:, 3] / 1e3))
        late = np.argsort(-ext_)[:5]
        print("   last CTAs (id, sm, entry, exit): " + ", ".join(f"({i}, {int(r[i, 2])}, {ent[i]:.2f}, {ext_[i]:.2f})" for i in late))
        tl.zero_()
        # a chain of back-to-back launches inside one CUDA graph, each with its own record buffer (the timeline pointer is
        # read at launch time and baked into the captured launch): the gap between the last exit of launch k and the first
        # entry of launch k + 1, and each launch's span, in %globaltimer time
        K = 8
        bufs = [torch.zeros(1024 + 4 * n, device=dev, dtype=torch.int64) for _ in range(K)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(K):
                L.load().tsc_debug_set_timeline(bufs[k].data_ptr())
                fn()
        L.load().tsc_debug_set_timeline(None)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        recs = [b[1024:].cpu().numpy().reshape(n, 4) for b in bufs]
        spans = [(r[:, 1].max() - r[:, 0].min()) / 1e3 for r in recs]
        gaps = [(recs[k + 1][:, 0].min() - recs[k][:, 1].max()) / 1e3 for k in range(K - 1)]
        print(f"   chain of {K} launches in a graph: {e0.elapsed_time(e1) * 1e3 / K:.2f} us per launch by events; "
              f"span first entry -> last exit per launch: " + " ".join(f"{v:.2f}" for v in spans))
        print("   gap last exit -> next launch's first entry [us]: " + " ".join(f"{v:.2f}" for v in gaps))


if __name__ == "__main__":
    main()
