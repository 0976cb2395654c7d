#!/bin/bash
# Round 2, call N (2 GPUs): the two-GPU training tests, the relight sweep tool at world size 2, the N=2 bench line.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "two_gpus" 2>&1 | tail -25 > gpurun_out/r2n_two_gpu_tests.log; cat gpurun_out/r2n_two_gpu_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/relight_sweep_bench.py --views 1 > gpurun_out/r2n_relight_n2.json 2> gpurun_out/r2n_relight_n2.err; tail -c 1500 gpurun_out/r2n_relight_n2.json; tail -5 gpurun_out/r2n_relight_n2.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err; tail -c 2500 gpurun_out/r2n_bench_n2.json; tail -5 gpurun_out/r2n_bench_n2.err
