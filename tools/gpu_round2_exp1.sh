#!/bin/bash
# experiment: k_bounce without its MLP (what does the SIMT part alone cost?)
set -x
mkdir -p gpurun_out
NMF_NVCC_EXTRA="-DNMF_BOUNCE_NO_MLP" timeout 900 python -m nmf_b200.build --force > gpurun_out/exp1_build.log 2>&1; tail -2 gpurun_out/exp1_build.log
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/exp1_bench.json 2> gpurun_out/exp1_bench.err; python tools/bench_phases.py gpurun_out/exp1_bench.json
