#!/bin/bash
# Round 2: compute-sanitizer racecheck over the shared-memory kernels added or changed this round.
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_train.py tests/test_gpu_mf_train.py -m gpu -q -x -k "overflowed or g40-24-True-None-0.0-True-fp32 or hand_over" > gpurun_out/r2_racecheck_train.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_racecheck_train.log | cut -c1-300
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_parity.py -m gpu -q -x -k "repack_kernels or device_resident or missing_and_empty" > gpurun_out/r2_racecheck_render.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_racecheck_render.log | cut -c1-300
