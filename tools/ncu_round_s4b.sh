#!/bin/bash
# DRAM traffic of every kernel of the eager cfg2 step (three metrics, one pass each) + the edge-case tests
python -m pytest tests/test_gpu_drivers.py -m gpu -q -rf -k "edge" 2>&1 | tail -30
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r1s4_traffic_step.csv python tools/step_once.py --steps 3 > gpurun_out/r1s4_ncu_traffic.log 2>&1
tail -n 2 gpurun_out/r1s4_ncu_traffic.log
ls -la gpurun_out/r1s4_traffic_step.csv
