#!/bin/bash
# Round 2, call B: reverse-pass parity (all cases), trainer, occupancy fix; timings + launch list of the training step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mf_train.py tests/test_gpu_train.py tests/test_gpu_plugins.py -m gpu -q -s 2>&1 > gpurun_out/r2b_tests_full.log; tail -30 gpurun_out/r2b_tests_full.log
timeout 600 python tools/mf_train_bench.py --retrace 1000,38000 > gpurun_out/r2b_mf_train_bench.json 2> gpurun_out/r2b_mf_train_bench.err; cat gpurun_out/r2b_mf_train_bench.json; tail -5 gpurun_out/r2b_mf_train_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_mf_train_launches.csv \
    python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2b_ncu.log 2>&1
grep -c k_ gpurun_out/r2b_mf_train_launches.csv
