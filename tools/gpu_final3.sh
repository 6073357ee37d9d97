#!/bin/bash
# final numbers of the session-4 build after the issue-loop change
tag=r1s4b
python bench.py > gpurun_out/${tag}_bench_cfg2_1gpu.json 2> gpurun_out/${tag}_bench_cfg2_1gpu.err
for w in cfg3 cfg4; do python bench.py --workload $w --steps 50 > gpurun_out/${tag}_bench_${w}_1gpu.json 2> /dev/null; done
python - <<PY
import json
for w in ("cfg2", "cfg3", "cfg4"):
    d = json.load(open("gpurun_out/${tag}_bench_%s_1gpu.json" % w))
    print(w, round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4))
PY
python tools/sweep.py --quick > gpurun_out/${tag}_sweep_quick.jsonl 2> gpurun_out/${tag}_sweep_quick.md
python tools/bench_eval.py --no-cpu > gpurun_out/${tag}_eval_gradnorm.jsonl 2> /dev/null; cut -c1-230 gpurun_out/${tag}_eval_gradnorm.jsonl
ncu --set full --clock-control none --import-source on -k regex:"osconv_tc_kernel" -c 3 -o gpurun_out/${tag}_conv_B1024 -f \
    python tools/prof_kernels.py --layer 1 --B 1024 --iters 1 > gpurun_out/${tag}_ncu_conv.log 2>&1
tail -n 3 gpurun_out/${tag}_ncu_conv.log
