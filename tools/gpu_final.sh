#!/bin/bash
# round-end check: smoke(), the whole GPU suite, the default bench line and the reference arm
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
python -m pytest tests -m gpu -q -x > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python bench.py > gpurun_out/r1s4_final_bench_cfg2_1gpu.json 2> gpurun_out/r1s4_final_bench_cfg2_1gpu.err; tail -2 gpurun_out/r1s4_final_bench_cfg2_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r1s4_final_bench_cfg2_1gpu.json')); print('cfg2', round(d['ms_per_step'],4), 'ms/step', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), d.get('cpu_baseline'), d['clocks'])"
