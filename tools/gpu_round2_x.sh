#!/bin/bash
# Round 2, call X: paired summed-area table (256-bit loads) in k_incoming, templated cluster k_select: suite, bench,
# training step, the demo with the mask threshold matched to the render alpha, racecheck of k_select.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2x_pytest_gpu.log; tail -5 gpurun_out/r2x_pytest_gpu.log
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 8 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; python tools/bench_phases.py gpurun_out/r2x_bench.json
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2x_mf_train.json 2>&1; tail -c 400 gpurun_out/r2x_mf_train.json
timeout 300 python tools/train_demo.py --iters 5000 --views 60 > gpurun_out/r2x_train_demo.json 2> gpurun_out/r2x_a.err; tail -c 200 gpurun_out/r2x_a.err
timeout 600 compute-sanitizer --tool racecheck --kernel-name kns=k_select python -m pytest tests/test_gpu_mf_train.py -m gpu -q -x -k "g40-24-True" > gpurun_out/r2x_racecheck.log 2>&1; tail -6 gpurun_out/r2x_racecheck.log
