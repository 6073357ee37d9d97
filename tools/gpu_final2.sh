#!/bin/bash
# final build: secondary workloads on one GPU, then the 2-GPU data-parallel line (run with gpurun --gpus 2)
for w in cfg3 cfg4; do
  python bench.py --workload $w --steps 50 > gpurun_out/r1s4_final_bench_${w}_1gpu.json 2> gpurun_out/r1s4_final_bench_${w}_1gpu.err
  python -c "
import json; d=json.load(open('gpurun_out/r1s4_final_bench_${w}_1gpu.json')); print('$w', round(d['ms_per_step'],4), 'ms/step', round(d['value']), d['unit'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r1s4_final_bench_cfg2_2gpu.json 2> gpurun_out/r1s4_final_bench_cfg2_2gpu.err
tail -2 gpurun_out/r1s4_final_bench_cfg2_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r1s4_final_bench_cfg2_2gpu.json').read().strip().splitlines()[-1]); print('cfg2 N=2', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'e2e', round(d['e2e']['value']))"
python bench.py --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-400
