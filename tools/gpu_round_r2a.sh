#!/bin/bash
# Round 2, call A: the whole GPU suite (incl. the new full-size parity tests), then the default bench line.
tag=${1:-r2a}
python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log
python bench.py --steps 50 > gpurun_out/${tag}_bench_cfg2_1gpu.json 2> gpurun_out/${tag}_bench_cfg2_1gpu.err; tail -3 gpurun_out/${tag}_bench_cfg2_1gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_cfg2_1gpu.json"))
print(round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac(in step)", round(d["roofline"]["frac"], 4),
      "isolated", round(d["roofline"]["isolated"]["frac"], 4), "launches", d["gpu_launches"])
print("gpu eager", d.get("gpu_eager_baseline")); print("tf32 peak", d.get("tf32_peak_tflops"))
for r in d.get("roofline_hbm", []) + d.get("roofline_gram", []): print("  ", r["kernel"], r["shape"], r["us"], "us", r["achieved"], r["unit"], r["frac"])
for k, v in d["kernels"].items(): print("   ", k, v)
PY
