#!/bin/bash
# Round 2, call C: the second-generation conv kernel -- correctness (kernel / property / module tests), then A/B timing.
tag=${1:-r2c}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -x > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log
for B in 128 1024; do
  for layer in 1 3; do
    echo "== v2 layer $layer B $B"; timeout 120 python tools/prof_kernels.py --layer $layer --B $B --iters 10 2>&1 | grep -v wgrad | tail -5
    echo "== v1 layer $layer B $B"; TSC_CONV_V1=1 timeout 120 python tools/prof_kernels.py --layer $layer --B $B --iters 10 2>&1 | grep -v wgrad | tail -3
  done
done
