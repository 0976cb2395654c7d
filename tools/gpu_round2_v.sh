#!/bin/bash
# Round 2, call V: cluster k_select (suite + racecheck of one training step + timings), train_demo with occupancy log.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2v_pytest_gpu.log; tail -5 gpurun_out/r2v_pytest_gpu.log
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2v_mf_train.json 2>&1; tail -c 600 gpurun_out/r2v_mf_train.json
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python tools/bench_phases.py gpurun_out/r2v_bench.json
timeout 300 python tools/train_demo.py --iters 5000 --views 60 > gpurun_out/r2v_train_demo.json 2> gpurun_out/r2v_train_demo.err; tail -c 300 gpurun_out/r2v_train_demo.err
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=k_select python -m pytest tests/test_gpu_mf_train.py -m gpu -q -x -k "retrace and fp32" > gpurun_out/r2v_racecheck.log 2>&1; tail -8 gpurun_out/r2v_racecheck.log
