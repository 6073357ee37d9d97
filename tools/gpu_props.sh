#!/bin/bash
python -m pytest tests/test_gpu_properties.py -m gpu -q -rf 2>&1 | tail -60 > gpurun_out/props_tests.log; tail -50 gpurun_out/props_tests.log
