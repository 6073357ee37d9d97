#!/usr/bin/env python
"""A few EAGER training steps of one bench workload and nothing else -- the target of the ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/step_once.py`), so that
the list holds exactly the launches of the step (bench.py adds isolated roofline loops and L2 flushes).
    python tools/step_once.py [--workload cfg2|cfg3|cfg4] [--steps 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                                # noqa: E402
import feature_level_style_transfer_for_tsc_b200 as T                      # noqa: E402
from feature_level_style_transfer_for_tsc_b200 import train_step as TS     # noqa: E402
from oracle import os_cnn as O                                              # noqa: E402  (synthetic input generator only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    T._lib.load()
    torch.manual_seed(0)
    if a.workload == "cfg3":
        c = bench.CFG3
        model = TS.MultiSourceModelSet(c["target"], c["sources"]).cuda()
        batches = [O.synthetic_batch(c["B"], *c["target"], 0)] + [O.synthetic_batch(c["B"], C, Ln, K, 1 + i)
                                                                  for i, (C, Ln, K) in enumerate(c["sources"])]
        ins = [t.cuda() for xb in batches for t in xb]
    elif a.workload == "cfg4":
        c = bench.CFG4
        model = TS.SingleDomainModelSet(c["C"], c["L"], c["K"]).cuda()
        ins = [t.cuda() for t in O.synthetic_batch(c["B"], c["C"], c["L"], c["K"], 0)]
    else:
        c = bench.CFG
        model = TS.StyleTransferModelSet(c["C"], c["L"], c["K"], c["C"], c["L"], c["K"]).cuda()
        ins = [t.cuda() for d in (0, 1) for t in O.synthetic_batch(c["B"], c["C"], c["L"], c["K"], d)]
    trainer = TS.Trainer(model, bench.STYLE_WEIGHT, use_graph=False)
    for _ in range(a.steps):
        loss = trainer.step(*ins)
    torch.cuda.synchronize()
    print(f"{a.workload}: {a.steps} eager steps, last loss {float(loss):.4f}")


if __name__ == "__main__":
    main()
