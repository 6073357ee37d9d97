#!/bin/bash
# bench.py at N GPUs of this box for the given workloads (run under `gpurun --gpus N`): one JSON line per run.
N=${1:-2}; tag=${2:-scale}
for w in ${WORKLOADS:-cfg2}; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --workload $w --no-cpu-baseline --steps 50 > gpurun_out/${tag}_${w}_n1.json 2> gpurun_out/${tag}_${w}_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 50 \
      > gpurun_out/${tag}_${w}_n$N.json 2> gpurun_out/${tag}_${w}_n$N.err
  fi
  tail -2 gpurun_out/${tag}_${w}_n$N.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${w}_n$N.json").read().strip().splitlines()[-1])
    print("$w N=$N", round(d["ms_per_step"], 4), "ms/step", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("no line:", e)
PY
done
