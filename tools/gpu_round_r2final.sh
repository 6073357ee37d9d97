#!/bin/bash
# Final evidence of round 2 (one B200): GPU suite, bench lines (cfg2 default, cfg3, cfg4), launch list + DRAM traffic of the
# eager cfg2 step, full ncu capture of the conv kernel at B=1024, cfg5 sweep, evaluation-path numbers.
tag=${1:-r2f}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 400 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; tail -2 gpurun_out/${tag}_bench_cfg2.err
for w in cfg3 cfg4; do
  timeout 300 python bench.py --workload $w --no-extra > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; tail -2 gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json
for w in ("cfg2", "cfg3", "cfg4"):
    try:
        d = json.load(open("gpurun_out/${tag}_bench_%s.json" % w))
        print(w, round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"])
        if w == "cfg2":
            for k, v in d["kernels"].items():
                print("   ", k, v)
    except Exception as e:
        print(w, "no line:", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${tag}_traffic_step.csv python tools/step_once.py --steps 3 > gpurun_out/${tag}_ncu_step.log 2>&1; tail -n 1 gpurun_out/${tag}_ncu_step.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"osconv2" -s 6 -c 4 -o gpurun_out/${tag}_conv2_B1024 -f \
    python tools/prof_kernels.py --layer 1 --B 1024 --iters 2 > gpurun_out/${tag}_ncu_conv2.log 2>&1; tail -n 2 gpurun_out/${tag}_ncu_conv2.log
timeout 400 python tools/sweep.py > gpurun_out/${tag}_sweep.jsonl 2> gpurun_out/${tag}_sweep.md; tail -n 3 gpurun_out/${tag}_sweep.md
timeout 300 python tools/bench_eval.py --no-cpu > gpurun_out/${tag}_eval.jsonl 2> gpurun_out/${tag}_eval.err; tail -n 2 gpurun_out/${tag}_eval.err; cat gpurun_out/${tag}_eval.jsonl | cut -c1-300
