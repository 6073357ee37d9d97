#!/bin/bash
# conv2 on CTA pairs: correctness first (bounded), then timing
tag=${1:-r2e}
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/${tag}_tests.log 2>&1; tail -15 gpurun_out/${tag}_tests.log
timeout 300 python -m pytest tests/test_gpu_properties.py -q -x > gpurun_out/${tag}_tests2.log 2>&1; tail -5 gpurun_out/${tag}_tests2.log
for B in 128 1024; do for layer in 1 3; do
    echo "== layer $layer B $B"; timeout 120 python tools/prof_kernels.py --layer $layer --B $B --iters 10 2>&1 | grep -v wgrad | tail -4
done; done
