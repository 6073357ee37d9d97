#!/bin/bash
# conv2: one-CTA (default) vs CTA-pair (TSC_CONV_PAIR=1) -- parity, then timing
for mode in 0 1; do
  echo "#### TSC_CONV_PAIR=$mode"
  TSC_CONV_PAIR=$mode timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -x 2>&1 | tail -2
  for B in 128 1024; do for layer in 1 3; do
    echo "== layer $layer B $B"; TSC_CONV_PAIR=$mode timeout 120 python tools/prof_kernels.py --layer $layer --B $B --iters 10 2>&1 | grep -v wgrad | grep "fwd \|dgrad \|timeline"
  done; done
done
