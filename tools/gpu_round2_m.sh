#!/bin/bash
# Round 2, call M: k_incoming with pipelined ray records + parallel tile descriptors; k_shade round-loop unroll experiment
# (rebuilt on the box with NMF_NVCC_EXTRA).
set -x
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3"
timeout 300 $B > gpurun_out/r2m_bench_base.json 2> gpurun_out/r2m_bench_base.err; python tools/bench_phases.py gpurun_out/r2m_bench_base.json
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for U in 2 4; do
  NMF_NVCC_EXTRA="-DNMF_SHADE_UNROLL=$U" timeout 900 python -m nmf_b200.build --force > gpurun_out/r2m_build_u$U.log 2>&1
  grep -A1 "k_shadeILi0ELi0" gpurun_out/r2m_build_u$U.log | head -4
  timeout 300 $B > gpurun_out/r2m_bench_u$U.json 2> gpurun_out/r2m_bench_u$U.err; python tools/bench_phases.py gpurun_out/r2m_bench_u$U.json
done
