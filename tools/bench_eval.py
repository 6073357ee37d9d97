#!/usr/bin/env python
"""Measurement of the 8f rows (evaluation / voting path, GradNorm joint-stage driver) on one B200.

    python tools/bench_eval.py [--steps 50] [--warmup 5] > profiles/<round>_eval_gradnorm.jsonl

One JSON line per measurement, CUDA-event timing on the launching stream, L2 flushed (256 MiB write) between timed
iterations, the CPU oracle (reference arithmetic, torch CPU kernels on all host cores) timed beside it on a bounded sample.
  * eval forward (extractor + classifier, eval-mode BatchNorm, no autograd) at cfg2's shape for B = 128 and 1024:
    the inference path (BatchNorm / ReLU / shortcut / pooling in the convolution epilogue) against the training kernels
    run forward-only, series/s and launches per pass;
  * the multi-source vote (3 models, N test series) and the per-class precision kernel;
  * one GradNorm joint-stage step (train_and_test.py:646-766 on cfg2-shaped modules) against the plain step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def timed(fn, steps, warmup, flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        flush.fill_(1.0)                              # L2 flush outside the events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import feature_level_style_transfer_for_tsc_b200 as T
    from feature_level_style_transfer_for_tsc_b200 import functional as TF
    from feature_level_style_transfer_for_tsc_b200 import grad_norm as D
    from feature_level_style_transfer_for_tsc_b200 import multi_source_voting as MV
    from feature_level_style_transfer_for_tsc_b200 import utils as U
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res
    from oracle import grad_norm as GN
    from oracle import os_cnn as O
    from oracle import voting as V
    T._lib.load()
    T.set_engine("tcgen05")
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda:0")
    flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)
    C, Ln, K = 9, 128, 6
    lpl_e, lpl_c = O.trainer_layer_lists(C, Ln)
    torch.manual_seed(0)
    fe, cl = OS_CNN_res(lpl_e).cuda(), OS_CNN(lpl_c, K).cuda()
    cores = os.cpu_count() or 1
    lines = []

    # ---- eval forward ---------------------------------------------------------------------------------------------
    for B in (128, 1024):
        x, y = O.synthetic_batch(B, C, Ln, K, 0)
        xd = x.cuda()
        fe.eval(); cl.eval()

        def fwd():
            with torch.no_grad():
                return cl(fe(xd))[0]

        ms_inf = timed(fwd, args.steps, args.warmup, flush)
        TF.INFERENCE_PATH = False
        ms_train_kernels = timed(fwd, args.steps, args.warmup, flush)
        TF.INFERENCE_PATH = True
        gp = U.GraphedPredictor([fe, cl], enabled=True)
        ms_graph = timed(lambda: gp(xd), args.steps, args.warmup, flush)
        # e2e: host batch -> device, forward, argmax + counts, one 4-byte read
        xh = x.pin_memory(); yh = y.pin_memory()

        def e2e():
            with torch.no_grad():
                lg = gp(xh.cuda(non_blocking=True))
                _, counts, _ = T.ops.class_precision(lg.contiguous(), yh.cuda(non_blocking=True))
                return int(counts[1].sum().item())

        ms_e2e = timed(e2e, args.steps, args.warmup, flush)
        flops = 2.0 * B * Ln * (sum(O.live_macs_per_position(l) for l in lpl_e) + C * O.feature_channels(lpl_e)
                                + sum(O.live_macs_per_position(l) for l in lpl_c))
        rec = dict(metric="OS-CNN eval forward samples/sec", unit="samples/s", workload=f"cfg2 shape, B={B}, C={C}, L={Ln}",
                   value=B / ms_graph * 1e3, ms=ms_graph, eager_ms=ms_inf, training_kernels_forward_only_eager_ms=ms_train_kernels,
                   cuda_graph=True,
                   e2e=dict(value=B / ms_e2e * 1e3, ms=ms_e2e, h2d_bytes=int(x.numel() * 4 + y.numel() * 8), d2h_bytes=4),
                   algorithmic_tflops=flops / ms_graph * 1e-9, dtype="bf16", l2="flushed between iterations")
        if not args.no_cpu and B == 128:
            torch.set_num_threads(cores)
            sd_fe = {k: v.detach().cpu().clone() for k, v in fe.state_dict().items()}
            sd_cl = {k: v.detach().cpu().clone() for k, v in cl.state_dict().items()}
            with torch.no_grad():
                O.classifier_forward(sd_cl, lpl_c, O.extractor_forward(sd_fe, lpl_e, x, training=False), training=False)
                t0 = time.perf_counter()
                n = 5
                for _ in range(n):
                    O.classifier_forward(sd_cl, lpl_c, O.extractor_forward(sd_fe, lpl_e, x, training=False), training=False)
                dt = (time.perf_counter() - t0) / n
            rec["cpu_baseline"] = dict(value=B / dt, unit="samples/s", cores=cores, kind="port", sample=f"{n} eval passes of B={B}")
        lines.append(rec)

    # ---- vote -----------------------------------------------------------------------------------------------------
    M, N = 3, 8192
    g = torch.Generator().manual_seed(1)
    tr = torch.randn(M, N, K, generator=g); te = torch.randn(M, N, K, generator=g)
    lab = torch.randint(0, K, (N,), generator=g)
    trd, ted, labd = tr.cuda(), te.cuda(), lab.cuda()

    def vote():
        precs = [T.ops.class_precision(trd[m], labd)[2] for m in range(M)]
        return MV.entropy_vote([ted[m] for m in range(M)], precs)

    ms_vote = timed(vote, args.steps, args.warmup, flush)
    rec = dict(metric="multi-source vote series/sec", unit="samples/s", workload=f"M={M} models, N={N} series, K={K} classes",
               value=N / ms_vote * 1e3, ms=ms_vote, launches=M + 1)
    if not args.no_cpu:
        t0 = time.perf_counter()
        precs = [V.class_precision(tr[m].numpy(), lab.numpy(), K) for m in range(M)]
        V.entropy_vote([te[m].numpy() for m in range(M)], V.normalized_weights(precs))
        rec["cpu_baseline"] = dict(value=N / (time.perf_counter() - t0), unit="samples/s", cores=1, kind="port",
                                   sample="one vote (the reference's per-row Python loop)")
    lines.append(rec)

    # ---- GradNorm joint-stage step ----------------------------------------------------------------------------------
    B = 128
    torch.manual_seed(0)
    mods = (OS_CNN_res(lpl_e).cuda(), OS_CNN(lpl_c, K).cuda(), OS_CNN_res(lpl_e).cuda(), OS_CNN(lpl_c, K).cuda())
    for m in mods:
        m.train()
    opts = [torch.optim.RMSprop(m.parameters(), lr=lr) for m, lr in zip(mods, (0.001, 0.003, 0.001, 0.003))]
    drv = D.JointStageDriver(mods[0].return_last_layer(), mods[2].return_last_layer(), opts)
    xt, yt = O.synthetic_batch(B, C, Ln, K, 0)
    xs, ys = O.synthetic_batch(B, C, Ln, K, 1)
    xt, yt, xs, ys = xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda()

    def gn_step():
        losses = GN.named_losses(mods, xt, yt, xs, ys, 1.0, adain=TF.adain, gram_style_loss=TF.gram_style_loss)
        drv.step(losses, 0)

    def plain_step():
        losses = GN.named_losses(mods, xt, yt, xs, ys, 1.0, adain=TF.adain, gram_style_loss=TF.gram_style_loss)
        for o in opts:
            o.zero_grad()
        sum(losses.values()).backward()
        for o in opts:
            o.step()

    ms_gn = timed(gn_step, max(10, args.steps // 2), args.warmup, flush)
    ms_plain = timed(plain_step, max(10, args.steps // 2), args.warmup, flush)
    rec = dict(metric="GradNorm joint-stage step", unit="samples/s", workload=f"cfg2 shape, B={B} per domain, eager (no CUDA graph), "
               "5 balanced losses: 5 extra backward passes with dense weight gradients over the last block",
               value=2 * B / ms_gn * 1e3, ms=ms_gn, plain_step_ms=ms_plain, gradnorm_overhead=ms_gn / ms_plain)
    # the same step as ONE CUDA graph (grad_norm.GraphedJointStage: forward, norm passes, device-side GradNorm, backward,
    # capturable optimizers); a failure here is reported, it does not take the other lines down
    try:
        torch.manual_seed(0)
        mods_g = (OS_CNN_res(lpl_e).cuda(), OS_CNN(lpl_c, K).cuda(), OS_CNN_res(lpl_e).cuda(), OS_CNN(lpl_c, K).cuda())
        for m in mods_g:
            m.train()
        opts_g = [torch.optim.RMSprop(m.parameters(), lr=lr, capturable=True) for m, lr in zip(mods_g, (0.001, 0.003, 0.001, 0.003))]
        drv_g = D.JointStageDriver(mods_g[0].return_last_layer(), mods_g[2].return_last_layer(), opts_g, capturable=True)
        stage = D.GraphedJointStage(drv_g, lambda a, b, c, d: GN.named_losses(mods_g, a, b, c, d, 1.0, adain=TF.adain,
                                                                             gram_style_loss=TF.gram_style_loss))
        ms_graph = timed(lambda: stage.step(xt, yt, xs, ys, cur_epoch=0), max(10, args.steps // 2), args.warmup, flush)
        rec.update(cuda_graph_ms=ms_graph, cuda_graph_value=2 * B / ms_graph * 1e3, cuda_graph_speedup=ms_gn / ms_graph)
    except Exception as e:                                  # noqa: BLE001
        rec.update(cuda_graph_error=repr(e)[:300])
    lines.append(rec)
    for r in lines:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
