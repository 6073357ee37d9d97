#!/bin/bash
# Round-2 conv experiments (run under gpurun on one B200; every command has its own timeout).  Results:
# profiles/r2_conv2_issue_bisection.md, profiles/r2_conv2_dual.md.
#   gpu_round_r2_conv.sh parity    kernel + property tests for the one-CTA and the CTA-pair variant, then timing of both
#   gpu_round_r2_conv.sh bisect    the TSC_C2_DEBUG knob matrix on the 72->228 bank (timing experiments, garbage results)
#   gpu_round_r2_conv.sh dual      full-size parity, then persistent CTAs with / without tile pairs per pass
mode=${1:-parity}
prof() { timeout 120 python tools/prof_kernels.py "$@" 2>&1 | grep -v wgrad | grep "fwd \|dgrad "; }
case $mode in
parity)
  for pair in 0 1; do
    echo "#### TSC_CONV_PAIR=$pair"
    TSC_CONV_PAIR=$pair timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -x 2>&1 | tail -2
    for B in 128 1024; do for layer in 1 2 3; do echo "== layer $layer B $B"; TSC_CONV_PAIR=$pair prof --layer $layer --B $B --iters 10; done; done
  done ;;
bisect)
  # 1 no copies | 16 same A rows | 32 same B rows | 64 N = np everywhere | 128 N = 112 everywhere | 256 accumulator poller sleeps
  # 512 producer sleeps | 1024 no weight pipeline at all | 2048 the micro-benchmark's loop of bare MMAs inside the kernel
  for dbg in 0 1 17 33 49 113 177 256 512 1024 3072; do
    echo "== TSC_C2_DEBUG=$dbg"
    for B in 128 1024; do TSC_C2_DEBUG=$dbg prof --layer 1 --B $B --iters 10; done
  done
  timeout 120 python tools/prof_kernels.py --layer 1 --B 128 --iters 5 --stages 2>&1 | grep -v wgrad | head -60 ;;
dual)
  timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py tests/test_gpu_fullsize.py tests/test_gpu_drivers.py -q -x 2>&1 | tail -4
  for dual in 1 0; do
    echo "#### TSC_C2_DUAL=$dual"
    for layer in 1 2 3; do TSC_C2_DUAL=$dual prof --layer $layer --B 1024 --iters 10; done
    TSC_C2_DUAL=$dual prof --layer 1 --B 4096 --iters 5
    TSC_C2_DUAL=$dual prof --C 3 --L 1024 --layer 1 --B 256 --iters 5
  done ;;
esac
