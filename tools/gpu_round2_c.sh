#!/bin/bash
# Round 2, call C: full GPU suite after the test fixes + occupancy linspace change.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2c_pytest_gpu.log; tail -15 gpurun_out/r2c_pytest_gpu.log
