#!/bin/bash
# 2 GPUs: does capping NCCL's CTAs make the early (overlapped) classifier slices pay?
tag=${1:-r2r}
i=0
for cfg in "1 4" "1 8" "0 8" "0 0"; do
  set -- $cfg; ov=$1; ctas=$2; i=$((i+1))
  if [ "$ctas" != "0" ]; then export NCCL_MAX_CTAS=$ctas; else unset NCCL_MAX_CTAS; fi
  TSC_DP_OVERLAP=$ov timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 2 --steps 100 --warmup 5 --no-extra > gpurun_out/${tag}_ov${ov}_ctas${ctas}.json 2> gpurun_out/${tag}_ov${ov}_ctas${ctas}.err; echo "overlap=$ov NCCL_MAX_CTAS=$ctas rc=$?"
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/${tag}_ov${ov}_ctas${ctas}.json | head -2
done
