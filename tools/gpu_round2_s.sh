#!/bin/bash
# Round 2, call S: suite with the segmented environment scans and the one-launch gradient hand-over; iteration breakdown.
set -x
mkdir -p gpurun_out
timeout 300 python tools/mf_iter_bench.py --steps 20 > gpurun_out/r2s_iter.json 2> gpurun_out/r2s_iter.err; cat gpurun_out/r2s_iter.json; tail -3 gpurun_out/r2s_iter.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2s_iter_launches.csv python tools/mf_iter_bench.py --steps 2 > gpurun_out/r2s_ncu.log 2>&1
