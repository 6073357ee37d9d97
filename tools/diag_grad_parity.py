"""Diagnostic (GPU): where does the gradient difference between the tcgen05 engine and the bf16-operand fp64 oracle come
from?  Runs the cfg2 step at B=128 on the three engines (tcgen05, simt_bf16 = same bf16 operands on CUDA cores, simt = fp32)
and prints pairwise per-tensor L2 differences for a few tensors, next to the two oracles."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import feature_level_style_transfer_for_tsc_b200 as T
from feature_level_style_transfer_for_tsc_b200.train_step import StyleTransferModelSet
from test_gpu_fullsize import _oracle_step, l2_rel

C, Ln, K, B = 9, 128, 6, 128
ref, g_exact, (xt, yt, xs, ys) = _oracle_step(C, Ln, K, B, 1.0, None)
emu, g_emu, _ = _oracle_step(C, Ln, K, B, 1.0, torch.bfloat16)
T._lib.load()
grads = {}
for eng in ("tcgen05", "simt_bf16", "simt"):
    T.set_engine(eng)
    torch.manual_seed(0)
    model = StyleTransferModelSet(C, Ln, K, C, Ln, K).cuda()
    model.train()
    out = model(xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda(), 1.0)
    out["loss"].backward()
    torch.cuda.synchronize()
    grads[eng] = {f"{g}.{k}": p.grad.detach().cpu() for g in ("fe_t", "cl_t", "fe_s", "du", "cl_s")
                  for k, p in getattr(model, g).named_parameters() if f"{g}.{k}" in g_exact}
keys = ["cl_t.net.2.conv1d.weight", "cl_t.net.1.conv1d.weight", "cl_t.net.0.conv1d.weight", "fe_t.net_1.net.net.1.conv1d.weight",
        "fe_t.net_1.net.net.0.conv1d.weight", "cl_t.net.1.bn.bias", "fe_s.net_1.net.net.1.bn.weight"]
rows = []
for k in keys:
    row = dict(tensor=k,
               tc_vs_simtbf16=l2_rel(grads["tcgen05"][k], grads["simt_bf16"][k]),
               tc_vs_emu=l2_rel(grads["tcgen05"][k], g_emu[k]),
               simtbf16_vs_emu=l2_rel(grads["simt_bf16"][k], g_emu[k]),
               simt_vs_exact=l2_rel(grads["simt"][k], g_exact[k]),
               emu_vs_exact=l2_rel(g_emu[k], g_exact[k]))
    rows.append(row)
    print(json.dumps(row))
