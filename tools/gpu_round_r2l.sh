#!/bin/bash
# (a) per-stage timeline of CTA 0 (72->228 bank, B=128)   (b) no weight pipeline at all: neither full-waits nor commits (1024)
timeout 120 python tools/prof_kernels.py --layer 1 --B 128 --iters 5 --stages 2>&1 | grep -v wgrad | head -80
for dbg in 1024 1200; do
  echo "== TSC_C2_DEBUG=$dbg"
  for B in 128 1024; do TSC_C2_DEBUG=$dbg timeout 120 python tools/prof_kernels.py --layer 1 --B $B --iters 10 2>&1 | grep "fwd \|dgrad \|timeline\[fwd\|timeline\[dgrad"; done
done
