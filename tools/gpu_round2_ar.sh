#!/bin/bash
# Round 2, call AR: two-level enumeration (8-step blocks skipped through the coarse occupancy field) in k_march pass A.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2as_pytest_gpu.log; tail -8 gpurun_out/r2as_pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2as_bench.json 2> gpurun_out/r2as_bench.err; python tools/bench_phases.py gpurun_out/r2as_bench.json; tail -2 gpurun_out/r2as_bench.err
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2as_mf_train.json 2>&1; tail -c 160 gpurun_out/r2as_mf_train.json
