#!/bin/bash
# Capture G: GPU suite, bench line, reference arm, ncu launch lists (render + training loop), full ncu of the optimiser kernels.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 400 gpurun_out/bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
grep -c k_ gpurun_out/launches.csv
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/train_launches.csv \
    python tools/train_bench.py --steps 2 --iters 5 > gpurun_out/ncu_train.log 2>&1
grep -c k_ gpurun_out/train_launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_adam|k_sq_norm|k_l1_reg' -c 4 \
    -o gpurun_out/prof_adam python tools/train_bench.py --steps 1 --iters 4 > gpurun_out/ncu_adam.log 2>&1
ls -la gpurun_out | tail -8
