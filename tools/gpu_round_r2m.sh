#!/bin/bash
# 2048(+1024: producer idle): the micro-benchmark's MMA loop inside conv2;  4096: table entry of the next run prefetched
for dbg in 3072 4096 4097; do
  echo "== TSC_C2_DEBUG=$dbg"
  for B in 128 1024; do TSC_C2_DEBUG=$dbg timeout 120 python tools/prof_kernels.py --layer 1 --B $B --iters 10 2>&1 | grep "fwd \|dgrad \|timeline\[fwd\|timeline\[dgrad"; done
done
