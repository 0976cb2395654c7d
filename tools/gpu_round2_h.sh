#!/bin/bash
# Round 2, call H: tcgen05 BRDF-MLP backward (bf16 operands, MN-major descriptors, TMEM-resident weight gradients):
# gradient parity of the f16 cases first (bounded by timeout: a wrong descriptor must not hang the box), then the suite + timing.
set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_mf_train.py -m gpu -q -s -x -k "f16 or accumulates" 2>&1 > gpurun_out/r2h_tc_bwd.log; tail -25 gpurun_out/r2h_tc_bwd.log | cut -c1-1500
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2h_pytest_gpu.log; tail -8 gpurun_out/r2h_pytest_gpu.log
timeout 600 python tools/mf_train_bench.py --retrace 1000,38000 > gpurun_out/r2h_mf_train_bench.json 2> gpurun_out/r2h_mf_train_bench.err; cut -c1-700 gpurun_out/r2h_mf_train_bench.json; tail -3 gpurun_out/r2h_mf_train_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_mf_train_launches.csv \
    python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2h_ncu.log 2>&1
