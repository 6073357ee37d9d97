#!/bin/bash
# 2 GPUs: bench line with the in-process NVML clock sampler
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_bench_2gpu.json | head -2; grep -o '"clocks": {[^}]*}' gpurun_out/r2v_bench_2gpu.json; grep -o "step_ms_spread.*" gpurun_out/r2v_bench_2gpu.json | cut -c1-260
