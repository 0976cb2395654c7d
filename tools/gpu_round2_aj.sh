#!/bin/bash
# Round 2, call AJ: experiment -- gather and shade phases of k_shade as two kernels (NMF_SHADE_SPLIT=1) against the fused kernel.
set -x
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3"
timeout 300 $B > gpurun_out/r2aj_bench_fused.json 2> gpurun_out/r2aj_a.err; python tools/bench_phases.py gpurun_out/r2aj_bench_fused.json
NMF_SHADE_SPLIT=1 timeout 300 $B > gpurun_out/r2aj_bench_split.json 2> gpurun_out/r2aj_b.err; python tools/bench_phases.py gpurun_out/r2aj_bench_split.json
NMF_SHADE_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -3
