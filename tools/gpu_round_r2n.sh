#!/bin/bash
# stage-structured issue loop (one elect per stage, runs inside): parity for both variants, then timing
for mode in 0 1; do
  echo "#### TSC_CONV_PAIR=$mode"
  TSC_CONV_PAIR=$mode timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -x 2>&1 | tail -2
  for B in 128 1024; do for layer in 1 2 3; do
    echo "== layer $layer B $B"; TSC_CONV_PAIR=$mode timeout 120 python tools/prof_kernels.py --layer $layer --B $B --iters 10 2>&1 | grep -v wgrad | grep "fwd \|dgrad \|timeline"
  done; done
done
echo "#### no weight pipeline (1024) / no copies (1)"
for dbg in 1024 1; do for B in 128 1024; do TSC_C2_DEBUG=$dbg timeout 120 python tools/prof_kernels.py --layer 1 --B $B --iters 10 2>&1 | grep "fwd \|dgrad "; done; done
