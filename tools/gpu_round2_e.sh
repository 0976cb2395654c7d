#!/bin/bash
# Round 2, call E (2 GPUs): the two-GPU training tests (plain + microfacet), the N=2 bench line with strong scaling and the
# sharded training iteration.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "two_gpus" 2>&1 | tail -20 > gpurun_out/r2e_two_gpu_tests.log; cat gpurun_out/r2e_two_gpu_tests.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; tail -c 2500 gpurun_out/r2e_bench_n2.json; tail -5 gpurun_out/r2e_bench_n2.err
