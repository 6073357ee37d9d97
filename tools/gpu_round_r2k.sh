#!/bin/bash
# Do the spinning waiters (accumulator poller, weight producer) slow the MMA issuer?  256: poller sleeps, 512: producer sleeps
for dbg in 256 512 768 769; do
  echo "== TSC_C2_DEBUG=$dbg"
  for B in 128 1024; do TSC_C2_DEBUG=$dbg timeout 120 python tools/prof_kernels.py --layer 1 --B $B --iters 10 2>&1 | grep "fwd \|dgrad "; done
done
