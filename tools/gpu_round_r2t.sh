#!/bin/bash
# dual-tile pass of the persistent conv CTAs: parity (kernel, property and full-size tests), then timing with / without it
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py tests/test_gpu_fullsize.py tests/test_gpu_drivers.py -q -x 2>&1 | tail -4
for dual in 1 0; do
  echo "#### TSC_C2_DUAL=$dual"
  for layer in 1 2 3; do
    TSC_C2_DUAL=$dual timeout 120 python tools/prof_kernels.py --layer $layer --B 1024 --iters 10 2>&1 | grep -v wgrad | grep "fwd \|dgrad "
  done
  TSC_C2_DUAL=$dual timeout 120 python tools/prof_kernels.py --layer 1 --B 4096 --iters 5 2>&1 | grep -v wgrad | grep "fwd \|dgrad "
  TSC_C2_DUAL=$dual timeout 120 python tools/prof_kernels.py --C 3 --L 1024 --layer 1 --B 256 --iters 5 2>&1 | grep -v wgrad | grep "fwd \|dgrad "
done
