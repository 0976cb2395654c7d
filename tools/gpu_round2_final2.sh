#!/bin/bash
# Round 2, final 2-GPU call: the two-GPU training tests (log kept under profiles/) and the N=2 bench line.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "two_gpus" 2>&1 | tail -12 > gpurun_out/r2g_two_gpu_tests.log; cat gpurun_out/r2g_two_gpu_tests.log | cut -c1-600
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; tail -c 1200 gpurun_out/r2g_bench_n2.json
