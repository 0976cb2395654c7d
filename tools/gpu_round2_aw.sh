#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mf_train.py tests/test_gpu_train.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2aw_mf_train.json 2>&1; tail -c 160 gpurun_out/r2aw_mf_train.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_mf_bounce|k_mf_tangent' -c 12 --csv --log-file gpurun_out/r2aw_launches.csv python tools/mf_train_bench.py --steps 1 --retrace 1000 > /dev/null 2>&1
grep "k_mf_" gpurun_out/r2aw_launches.csv | tail -3 | awk -F'","' '{print $5, $NF}' | cut -c1-80
