#!/bin/bash
# Round 2: compute-sanitizer memcheck over small cases that reach every kernel changed this round (render with retrace,
# microfacet training step + trainer iteration with the repack kernels, cluster k_select, tile descriptors, 256-bit records).
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_mf_train.py -m gpu -q -x -k "g40-24-True-None-0.0-True-fp32 or g40-24-False-None-0.0-True-f16 or edge or hand_over" > gpurun_out/r2_memcheck_train.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_memcheck_train.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_train.py tests/test_gpu_plugins.py -m gpu -q -x -k "overflowed or device_resident or repack_kernels" > gpurun_out/r2_memcheck_iter.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_memcheck_iter.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "missing_and_empty or overflow or matches_oracle" > gpurun_out/r2_memcheck_render.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_memcheck_render.log | cut -c1-300
