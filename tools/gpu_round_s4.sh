#!/bin/bash
# One GPU-box round of session 4 (run under gpurun): smoke(), the GPU test suite, the bench lines of the three workloads, the
# quick cfg5 sweep, the evaluation / GradNorm measurements and a full ncu capture of the conv kernel.  Outputs -> gpurun_out/.
tag=${1:-r1s4b}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench_cfg2_1gpu.json 2> gpurun_out/${tag}_bench_cfg2_1gpu.err
for w in cfg3 cfg4; do python bench.py --workload $w --steps 50 > gpurun_out/${tag}_bench_${w}_1gpu.json 2> /dev/null; done
python - <<PY
import json
for w in ("cfg2", "cfg3", "cfg4"):
    d = json.load(open("gpurun_out/${tag}_bench_%s_1gpu.json" % w))
    print(w, round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4))
PY
python tools/sweep.py --quick > gpurun_out/${tag}_sweep_quick.jsonl 2> gpurun_out/${tag}_sweep_quick.md
python tools/bench_eval.py > gpurun_out/${tag}_eval_gradnorm.jsonl 2> /dev/null; cut -c1-230 gpurun_out/${tag}_eval_gradnorm.jsonl
ncu --set full --clock-control none --import-source on -k regex:"osconv_tc_kernel" -c 3 -o gpurun_out/${tag}_conv_B1024 -f \
    python tools/prof_kernels.py --layer 1 --B 1024 --iters 1 > gpurun_out/${tag}_ncu_conv.log 2>&1
tail -n 3 gpurun_out/${tag}_ncu_conv.log
# 2-GPU line: gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#   --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5'
