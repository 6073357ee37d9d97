#!/bin/bash
# Conv-kernel experiment matrix (profiles/README.md): weight-ring depth / stage size / issue-only / stream-only.
out=${1:-gpurun_out/conv_knobs.txt}
: > $out
run() {
  echo "=== $1 ===" >> $out
  for B in 128 1024; do
    env $1 python tools/prof_kernels.py --layer 1 --B $B --iters 10 2>&1 | grep -E "^(fwd|dgrad|timeline\[(fwd|dgrad))" >> $out
  done
}
for k in ${KNOBS:-"TSC_X=0" "TSC_CONV_DEBUG=2" "TSC_CONV_DEBUG=4" "TSC_CONV_DEBUG=6"}; do run "$k"; done
cat $out
