#!/bin/bash
# Round 2, call L: suite with the tile-descriptor walk (k_bounce / k_incoming) and the parallel threshold scan (k_select),
# the full 1-GPU bench line, the microfacet training-step bench, launch list of the training step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2l_pytest_gpu.log; tail -6 gpurun_out/r2l_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -c 1500 gpurun_out/r2l_bench.json; tail -3 gpurun_out/r2l_bench.err
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2l_mf_train.json 2>&1; tail -c 1500 gpurun_out/r2l_mf_train.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2l_ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 200 --csv --log-file gpurun_out/r2l_train_launches.csv \
    python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2l_ncu_train.log 2>&1
