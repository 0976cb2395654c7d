#!/bin/bash
# Round 2, call D: bench line with the new legs (L2 gather ceiling, e2e_plugin, sustained, train_step, reference_cuda),
# both reference arms.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err; tail -c 300 gpurun_out/r2d_bench_ref.json
timeout 600 python bench.py --impl reference-cuda --steps 3 --warmup 1 > gpurun_out/r2d_bench_refcuda.json 2> gpurun_out/r2d_bench_refcuda.err; tail -c 1200 gpurun_out/r2d_bench_refcuda.json; tail -5 gpurun_out/r2d_bench_refcuda.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 3000 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
