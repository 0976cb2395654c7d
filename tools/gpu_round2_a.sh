#!/bin/bash
# Round 2, call A: the microfacet reverse pass (new) + the established suite.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mf_train.py -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2a_mf_train.log; tail -40 gpurun_out/r2a_mf_train.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_mf_train.py 2>&1 | tail -15 > gpurun_out/r2a_pytest_gpu.log; tail -5 gpurun_out/r2a_pytest_gpu.log
