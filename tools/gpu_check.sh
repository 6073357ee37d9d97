#!/bin/bash
# One GPU-box round: the GPU test suite, then the default bench line (and optionally another workload) into gpurun_out/.
tag=${1:-check}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
for w in ${WORKLOADS:-cfg2}; do
  timeout 300 python bench.py --workload $w ${BENCH_ARGS:-} > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; tail -2 gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_$w.json"))
print("$w", round(d["ms_per_step"], 4), "ms/step", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "roofline frac", round(d["roofline"]["frac"], 4),
      "launches", d["gpu_launches"], d.get("cpu_baseline"))
PY
done
