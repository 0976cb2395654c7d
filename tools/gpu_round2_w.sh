#!/bin/bash
# Round 2, call W: train_demo with every occupancy rebuild recorded (compressed schedule vs the reference's absolute
# schedule), racecheck of the cluster k_select.
set -x
mkdir -p gpurun_out
timeout 300 python tools/train_demo.py --iters 5000 --views 60 > gpurun_out/r2w_train_demo_compressed.json 2> gpurun_out/r2w_a.err; tail -c 200 gpurun_out/r2w_a.err
timeout 300 python tools/train_demo.py --iters 6000 --views 60 --ups 500,1000,2000 --upd 2000,3000,4000 > gpurun_out/r2w_train_demo_refsched.json 2> gpurun_out/r2w_b.err; tail -c 200 gpurun_out/r2w_b.err
timeout 600 compute-sanitizer --tool racecheck --kernel-name regex:k_select python -m pytest tests/test_gpu_mf_train.py -m gpu -q -x -k "g40-24-True" > gpurun_out/r2w_racecheck.log 2>&1; tail -6 gpurun_out/r2w_racecheck.log
