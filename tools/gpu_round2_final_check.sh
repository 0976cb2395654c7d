#!/bin/bash
# Round 2, final sanity on 2 GPUs with the last build: smoke(), the N=2 bench line, the two-GPU tests.
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; tail -c 900 gpurun_out/r2j_bench_n2.json; tail -2 gpurun_out/r2j_bench_n2.err
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -k "two_gpus" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 300
