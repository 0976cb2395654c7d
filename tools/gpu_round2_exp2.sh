#!/bin/bash
# experiment: k_march with pass B (density of the valid samples) disabled: what does the dense-step enumeration alone cost?
set -x
mkdir -p gpurun_out
NMF_NVCC_EXTRA="-DNMF_MARCH_NO_PASS_B" timeout 900 python -m nmf_b200.build --force > gpurun_out/exp2_build.log 2>&1; tail -1 gpurun_out/exp2_build.log
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 5 --warmup 3 > gpurun_out/exp2_bench.json 2> gpurun_out/exp2_bench.err; python tools/bench_phases.py gpurun_out/exp2_bench.json; tail -2 gpurun_out/exp2_bench.err
