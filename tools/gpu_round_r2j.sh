#!/bin/bash
# In-kernel bisection of the MMA issue cost (conv2, 72->228 bank, garbage results): the weight stream is off (bit 1) in all runs
for dbg in 1 17 33 49 65 113 129 177; do
  echo "== TSC_C2_DEBUG=$dbg"
  for B in 128 1024; do TSC_C2_DEBUG=$dbg timeout 120 python tools/prof_kernels.py --layer 1 --B $B --iters 10 2>&1 | grep "fwd \|dgrad "; done
done
