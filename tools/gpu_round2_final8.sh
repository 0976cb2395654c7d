#!/bin/bash
# Round 2, final 8-GPU lines: bench --gpus 8 (weak + strong scaling + ray-sharded training iteration) and the relight sweep.
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err; tail -c 1500 gpurun_out/r2g_bench_n8.json; tail -3 gpurun_out/r2g_bench_n8.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 tools/relight_sweep_bench.py --views 4 > gpurun_out/r2g_relight_n8.json 2> gpurun_out/r2g_relight_n8.err; tail -c 700 gpurun_out/r2g_relight_n8.json
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n4.json 2> gpurun_out/r2g_bench_n4.err; tail -c 900 gpurun_out/r2g_bench_n4.json
