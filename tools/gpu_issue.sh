#!/bin/bash
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -m gpu -q -x -k "conv or integer or inference" 2>&1 | tail -3
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python tools/sweep.py --quick 2>&1 >/dev/null | grep -E "osconv" 
python bench.py --no-cpu-baseline --steps 100 > gpurun_out/issue_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/issue_bench.json')); print('cfg2', round(d['ms_per_step'],4), 'ms/step frac', round(d['roofline']['frac'],4)); print(d['kernels']['osconv[tc]'])"
