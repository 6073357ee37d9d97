#!/bin/bash
# full GPU suite + the default bench line with the stage-structured conv issue loop
tag=${1:-r2o}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; tail -2 gpurun_out/${tag}_bench_cfg2.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_cfg2.json"))
print(round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "iso frac", round(d["roofline"]["isolated"]["frac"], 4), "launches", d["gpu_launches"])
for k, v in d["kernels"].items():
    print("   ", k, v)
PY
