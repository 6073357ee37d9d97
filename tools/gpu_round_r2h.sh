#!/bin/bash
# Round-2 re-entry round: the whole GPU suite, the default bench line, the launch list + DRAM traffic of the eager
# cfg2 step, and a full ncu capture of the persistent conv kernel at B=1024 (layer 1 = the 72->228 bank).
tag=${1:-r2h}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; tail -2 gpurun_out/${tag}_bench_cfg2.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_cfg2.json"))
print(round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"])
for k, v in d["kernels"].items():
    print("   ", k, v)
print(d.get("gpu_eager_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${tag}_traffic_step.csv python tools/step_once.py --steps 3 > gpurun_out/${tag}_ncu_step.log 2>&1
tail -n 1 gpurun_out/${tag}_ncu_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"osconv2" -s 6 -c 4 -o gpurun_out/${tag}_conv2_B1024 -f \
    python tools/prof_kernels.py --layer 1 --B 1024 --iters 2 > gpurun_out/${tag}_ncu_conv2.log 2>&1
tail -n 4 gpurun_out/${tag}_ncu_conv2.log
ls -la gpurun_out/
