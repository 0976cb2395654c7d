#!/bin/bash
# Round 2, call G: suite after repack kernels / v4 normal atomics / skip of no-op survivors; training-step timing + launch list.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/r2g_pytest_gpu.log; tail -8 gpurun_out/r2g_pytest_gpu.log
timeout 600 python tools/mf_train_bench.py --retrace 1000,38000 > gpurun_out/r2g_mf_train_bench.json 2> gpurun_out/r2g_mf_train_bench.err; cat gpurun_out/r2g_mf_train_bench.json | cut -c1-700; tail -3 gpurun_out/r2g_mf_train_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_mf_train_launches.csv \
    python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2g_ncu.log 2>&1
grep -c k_ gpurun_out/r2g_mf_train_launches.csv
