#!/bin/bash
# lean inline-PTX issue loop of conv2: parity (kernel + property tests) for the one-CTA and the CTA-pair variant, then timing
for mode in 0 1; do
  echo "#### TSC_CONV_PAIR=$mode"
  TSC_CONV_PAIR=$mode timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -x 2>&1 | tail -2
  for B in 128 1024; do for layer in 1 2 3; do
    echo "== layer $layer B $B"; TSC_CONV_PAIR=$mode timeout 120 python tools/prof_kernels.py --layer $layer --B $B --iters 10 2>&1 | grep -v wgrad | grep "fwd \|dgrad \|timeline"
  done; done
done
echo "#### no-copy experiment (TSC_C2_DEBUG=1, garbage results): is the weight stream a bound?"
for B in 128 1024; do TSC_C2_DEBUG=1 timeout 120 python tools/prof_kernels.py --layer 1 --B $B --iters 10 2>&1 | grep "fwd \|dgrad "; done
