#!/bin/bash
# Round 2, call AT: summed-area-table boxes whose four corners share one 2 x 2 texel block load it once.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2at_pytest_gpu.log; tail -4 gpurun_out/r2at_pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err; python tools/bench_phases.py gpurun_out/r2at_bench.json; tail -2 gpurun_out/r2at_bench.err
