#!/bin/bash
timeout 300 python bench.py > gpurun_out/r2v_bench_1gpu.json 2> gpurun_out/r2v_bench_1gpu.err; echo "rc=$?"; tail -2 gpurun_out/r2v_bench_1gpu.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_bench_1gpu.json | head -2; grep -o '"clocks": {[^}]*}' gpurun_out/r2v_bench_1gpu.json; grep -o "step_ms_spread.*" gpurun_out/r2v_bench_1gpu.json | cut -c1-260
