#!/bin/bash
# conv2 experiments: ring depth, stage size, no-copy, persistent grid size
run() { echo "== $1"; env $1 timeout 120 python tools/prof_kernels.py --layer 1 --B $2 --iters 10 2>&1 | grep -v wgrad | grep "fwd\|dgrad" ; }
run "X=0" 128
run "TSC_C2_SMEM_FULL=1" 128
run "TSC_C2_SMEM_FULL=1 TSC_C2_STAGE_KB=16" 128
run "TSC_C2_DEBUG=1" 128
run "TSC_C2_SMEM_FULL=1 TSC_C2_DEBUG=1" 128
run "X=0" 1024
run "TSC_C2_DEBUG=1" 1024
run "TSC_C2_GRID=74" 1024
run "TSC_C2_GRID=74 TSC_C2_DEBUG=1" 1024
run "TSC_C2_STAGE_KB=16" 1024
