#!/bin/bash
# Round 2, last 1-GPU measurement of the final build: suite, full bench line, training step / iteration, launch list of the step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2k3_pytest_gpu.log; tail -4 gpurun_out/r2k3_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2k3_bench.json 2> gpurun_out/r2k3_bench.err; python tools/bench_phases.py gpurun_out/r2k3_bench.json; tail -2 gpurun_out/r2k3_bench.err
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2k3_mf_train.json 2>&1; tail -c 200 gpurun_out/r2k3_mf_train.json
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 38000 > gpurun_out/r2k3_mf_train_38k.json 2>&1; tail -c 200 gpurun_out/r2k3_mf_train_38k.json
timeout 300 python tools/mf_iter_bench.py --steps 20 > gpurun_out/r2k3_iter.json 2> gpurun_out/r2k3_iter.err; cat gpurun_out/r2k3_iter.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 200 --csv --log-file gpurun_out/r2k3_train_launches.csv python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2k3_ncu_train.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k3_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2k3_ncu_launches.log 2>&1
