#!/usr/bin/env python
"""A few EAGER evaluation passes (extractor + classifier in eval mode, then the vote kernels) and nothing else -- the target
of the ncu launch list / full capture of the inference path.
    python tools/eval_once.py [--B 1024] [--passes 2]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_level_style_transfer_for_tsc_b200 as T                                      # noqa: E402
from feature_level_style_transfer_for_tsc_b200 import multi_source_voting as MV           # noqa: E402
from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res    # noqa: E402
from oracle import os_cnn as O                                                              # noqa: E402  (layer lists, inputs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--passes", type=int, default=2)
    a = ap.parse_args()
    T._lib.load()
    T.set_engine("tcgen05")
    C, Ln, K = 9, 128, 6
    lpl_e, lpl_c = O.trainer_layer_lists(C, Ln)
    torch.manual_seed(0)
    fe, cl = OS_CNN_res(lpl_e).cuda().eval(), OS_CNN(lpl_c, K).cuda().eval()
    x, y = O.synthetic_batch(a.B, C, Ln, K, 0)
    x, y = x.cuda(), y.cuda()
    with torch.no_grad():
        for _ in range(a.passes):
            logits = cl(fe(x))[0]
        prec = T.ops.class_precision(logits.contiguous(), y)[2]
        MV.entropy_vote([logits, logits, logits], [prec, prec, prec])
    torch.cuda.synchronize()
    print("ok", float(logits.abs().mean()))


if __name__ == "__main__":
    main()
