#!/bin/bash
# Round 2, call AK: SFU math for the feature noise.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2ak_pytest_gpu.log; tail -4 gpurun_out/r2ak_pytest_gpu.log
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; python tools/bench_phases.py gpurun_out/r2ak_bench.json
