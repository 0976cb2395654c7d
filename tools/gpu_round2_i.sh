#!/bin/bash
# Round 2, call I: tcgen05 reverse MLP, both operand-format variants (all-BF16 vs FP16 activations/weights + BF16 gradients)
set -x
mkdir -p gpurun_out
NMF_TC_BWD_MIXED=0 timeout 180 python -m pytest tests/test_gpu_mf_train.py -m gpu -q -s -k "f16" 2>&1 | grep "live_N\|passed\|failed" | cut -c1-1300 > gpurun_out/r2i_tc_bf16.log; cat gpurun_out/r2i_tc_bf16.log
NMF_TC_BWD_MIXED=1 timeout 180 python -m pytest tests/test_gpu_mf_train.py -m gpu -q -s -k "f16" 2>&1 | grep "live_N\|passed\|failed" | cut -c1-1300 > gpurun_out/r2i_tc_mixed.log; cat gpurun_out/r2i_tc_mixed.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2i_pytest_gpu.log; tail -6 gpurun_out/r2i_pytest_gpu.log
