#!/bin/bash
# First GPU call of the next round: gate the reverse-pass kernels written in round 1's last GPU seconds (they sort last
# in the suite), time them, and capture them once with ncu.  Everything else in the suite is the established parity set.
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_zz_env_bwd.py -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_reverse.log; tail -5 gpurun_out/pytest_reverse.log
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/reverse_bench.py > gpurun_out/reverse_bench.json 2> gpurun_out/reverse_bench.err; tail -c 1500 gpurun_out/reverse_bench.json; tail -3 gpurun_out/reverse_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_env_bwd|k_normals_bwd|k_heads_bwd' -c 12 \
    -o gpurun_out/prof_reverse python tools/reverse_bench.py --iters 1 > gpurun_out/ncu_reverse.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
ls -la gpurun_out | tail -6
