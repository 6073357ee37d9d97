#!/bin/bash
# 8 GPUs: the cfg2 bench line with the whole step (NCCL all-reduce + RMSprop inside) in one CUDA graph per rank
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 100 --warmup 5 --no-extra > gpurun_out/r2_bench_cfg2_8gpu_b.json 2> gpurun_out/r2_bench_cfg2_8gpu_b.err; echo "rc=$?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_cfg2_8gpu_b.json | head -2; grep -o '"value": [0-9.]*' gpurun_out/r2_bench_cfg2_8gpu_b.json | head -1
grep -o "step_ms_spread.*" gpurun_out/r2_bench_cfg2_8gpu_b.json | cut -c1-300
