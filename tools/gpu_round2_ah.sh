#!/bin/bash
# Round 2, call AH: one-wave grids for the reverse-pass kernels; experiment: two normal-gather groups in flight in k_mf_sample_bwd.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mf_train.py tests/test_gpu_train.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2ah_mf_train.json 2>&1; tail -c 200 gpurun_out/r2ah_mf_train.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 120 --csv --log-file gpurun_out/r2ah_train_launches.csv python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2ah_ncu.log 2>&1
NMF_NVCC_EXTRA="-DNMF_SB_UNROLL=2" timeout 900 python -m nmf_b200.build --force > gpurun_out/r2ah_build.log 2>&1; grep -A2 "k_mf_sample_bwdILi0" gpurun_out/r2ah_build.log | head -4
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2ah_mf_train_u2.json 2>&1; tail -c 200 gpurun_out/r2ah_mf_train_u2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_mf_sample' -c 20 --csv --log-file gpurun_out/r2ah_train_launches_u2.csv python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2ah_ncu2.log 2>&1
