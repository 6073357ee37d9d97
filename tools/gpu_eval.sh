#!/bin/bash
python -m pytest tests/test_gpu_drivers.py -m gpu -q -rf -k "inference" 2>&1 | tail -40 > gpurun_out/eval_tests.log; tail -30 gpurun_out/eval_tests.log
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_drivers.py::test_inference_path_equals_the_training_kernels_in_eval_mode > gpurun_out/eval_all_tests.log 2>&1; tail -3 gpurun_out/eval_all_tests.log
timeout 300 python tools/bench_eval.py > gpurun_out/r1s4_eval_gradnorm.jsonl 2> gpurun_out/r1s4_eval_gradnorm.err; tail -3 gpurun_out/r1s4_eval_gradnorm.err; cat gpurun_out/r1s4_eval_gradnorm.jsonl
