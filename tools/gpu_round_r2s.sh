#!/bin/bash
# short-series (wide-layer) route + module / kernel tests
timeout 300 python -m pytest tests/test_gpu_modules.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -15
