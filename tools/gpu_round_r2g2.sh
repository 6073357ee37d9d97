#!/bin/bash
# final build: default bench line + launch list / DRAM traffic of the eager cfg2 step
tag=${1:-r2g2}
timeout 400 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; tail -2 gpurun_out/${tag}_bench_cfg2.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_cfg2.json"))
print(round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "iso", round(d["roofline"]["isolated"]["frac"], 4), "launches", d["gpu_launches"], d["gpu_eager_baseline"]["ms_per_step"], d["cpu_baseline"]["value"])
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${tag}_traffic_step.csv python tools/step_once.py --steps 3 > gpurun_out/${tag}_ncu_step.log 2>&1; tail -n 1 gpurun_out/${tag}_ncu_step.log
