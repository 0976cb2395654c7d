#!/bin/bash
# Round 2, call AI: per-ray segmented sums for the auxiliary-map atomics of k_shade<0>.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2ai_pytest_gpu.log; tail -4 gpurun_out/r2ai_pytest_gpu.log
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err; python tools/bench_phases.py gpurun_out/r2ai_bench.json
