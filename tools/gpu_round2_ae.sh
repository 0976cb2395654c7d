#!/bin/bash
# Round 2, call AE: 256-bit line taps in the normal gather (k_shade): suite + bench phases + training step.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2ae_pytest_gpu.log; tail -4 gpurun_out/r2ae_pytest_gpu.log
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; python tools/bench_phases.py gpurun_out/r2ae_bench.json
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2ae_mf_train.json 2>&1; tail -c 300 gpurun_out/r2ae_mf_train.json
