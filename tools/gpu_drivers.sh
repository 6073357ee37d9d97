#!/bin/bash
# GPU round for the 8f rows: the new tests verbosely (all failures shown), then the whole GPU suite, then a short bench line.
python -m pytest tests/test_gpu_drivers.py -m gpu -q -rf 2>&1 | tail -80 > gpurun_out/drivers_tests.log; tail -60 gpurun_out/drivers_tests.log
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_drivers.py > gpurun_out/drivers_all_tests.log 2>&1; tail -3 gpurun_out/drivers_all_tests.log
python bench.py --no-cpu-baseline --steps 50 > gpurun_out/drivers_bench.json 2> gpurun_out/drivers_bench.err
python -c "
import json; d=json.load(open('gpurun_out/drivers_bench.json')); print('cfg2', round(d['ms_per_step'],4), 'ms/step')"
