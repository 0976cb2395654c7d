#!/bin/bash
# Round 2, call AV: environment scalars loaded once per kernel (NmfEnvDyn) instead of pointer selects into the parameter block.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2av_pytest_gpu.log; tail -3 gpurun_out/r2av_pytest_gpu.log | cut -c1-200
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2av_bench.json 2> gpurun_out/r2av_bench.err; python tools/bench_phases.py gpurun_out/r2av_bench.json
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2av_mf_train.json 2>&1; tail -c 160 gpurun_out/r2av_mf_train.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_mf_bounce|k_mf_tangent|k_mf_sec' -c 20 --csv --log-file gpurun_out/r2av_launches.csv python tools/mf_train_bench.py --steps 1 --retrace 1000 > /dev/null 2>&1
grep "k_mf_" gpurun_out/r2av_launches.csv | tail -5 | awk -F'","' '{print $5, $NF}' | cut -c1-80
timeout 300 python tools/mf_iter_bench.py --steps 20 2>/dev/null | tail -c 400
