#!/bin/bash
# Round-2 data-parallel measurements (run under `gpurun --gpus N`; every command has a SHORT timeout: a hung rank is charged
# N x the wall time).  Results: profiles/r2_dp_exchange.md, profiles/r2_dp_equality_2gpu.jsonl, profiles/r2_bench_cfg2_{2,8}gpu.json.
#   gpu_round_r2_dp.sh equality         tests/dp_worker.py on 2 GPUs (whole step in one graph, NCCL inside)
#   gpu_round_r2_dp.sh variants         2-GPU bench: exchange after backward vs early classifier slices, NCCL CTA caps
#   gpu_round_r2_dp.sh bench N          the N-GPU bench line
mode=${1:-equality}
run() { n=$1; port=$2; shift 2; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
case $mode in
equality)
  run 2 29517 tests/dp_worker.py > gpurun_out/r2_dp_equality.jsonl 2> gpurun_out/r2_dp_equality.err; echo "rc=$?"; cat gpurun_out/r2_dp_equality.jsonl ;;
variants)
  i=0
  for cfg in "0 0" "1 0" "1 8" "1 4" "0 8"; do
    set -- $cfg; ov=$1; ctas=$2; i=$((i+1))
    if [ "$ctas" != "0" ]; then export NCCL_MAX_CTAS=$ctas; else unset NCCL_MAX_CTAS; fi
    TSC_DP_OVERLAP=$ov run 2 2954$i bench.py --gpus 2 --steps 100 --warmup 5 --no-extra > gpurun_out/r2_dp_ov${ov}_ctas${ctas}.json 2> gpurun_out/r2_dp_ov${ov}_ctas${ctas}.err
    echo "overlap=$ov NCCL_MAX_CTAS=$ctas rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp_ov${ov}_ctas${ctas}.json | head -2
  done ;;
bench)
  n=${2:-2}
  run $n 29561 bench.py --gpus $n --steps 100 --warmup 5 --no-extra > gpurun_out/r2_bench_cfg2_${n}gpu.json 2> gpurun_out/r2_bench_cfg2_${n}gpu.err; echo "rc=$?"
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_cfg2_${n}gpu.json | head -2; grep -o "step_ms_spread.*" gpurun_out/r2_bench_cfg2_${n}gpu.json | cut -c1-260 ;;
esac
