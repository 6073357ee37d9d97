#!/bin/bash
# 2 GPUs, every command under a short timeout: the data-parallel equality worker (whole step in one graph, NCCL inside), then
# the 2-GPU bench line with and without the early (overlapped) classifier slices
tag=${1:-r2q}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dp_worker.py > gpurun_out/${tag}_dp.jsonl 2> gpurun_out/${tag}_dp.err; echo "dp worker rc=$?"; cat gpurun_out/${tag}_dp.jsonl; grep "rank 0" gpurun_out/${tag}_dp.err | tail -4
for ov in 1 0; do
  TSC_DP_OVERLAP=$ov timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$ov bench.py --gpus 2 --steps 100 --warmup 5 --no-extra > gpurun_out/${tag}_bench_2gpu_ov$ov.json 2> gpurun_out/${tag}_bench_2gpu_ov$ov.err; echo "bench overlap=$ov rc=$?"
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/${tag}_bench_2gpu_ov$ov.json | head -2
done
