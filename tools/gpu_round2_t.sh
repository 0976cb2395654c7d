#!/bin/bash
# Round 2, call T (8 GPUs): the 18-job relight sweep (BASELINE config #5) and the N=8 bench line (weak scaling + strong
# scaling of one image + the ray-sharded training iteration with the all-reduce broken out).
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 tools/relight_sweep_bench.py --views 4 > gpurun_out/r2t_relight_n8.json 2> gpurun_out/r2t_relight_n8.err; tail -c 1500 gpurun_out/r2t_relight_n8.json; tail -5 gpurun_out/r2t_relight_n8.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; tail -c 1800 gpurun_out/r2t_bench_n8.json; tail -5 gpurun_out/r2t_bench_n8.err
