#!/bin/bash
# Round 2, call O: where a MicrofacetTrainer iteration spends its time (wall clock per part + launch list).
set -x
mkdir -p gpurun_out
timeout 300 python tools/mf_iter_bench.py --steps 20 > gpurun_out/r2o_iter.json 2> gpurun_out/r2o_iter.err; cat gpurun_out/r2o_iter.json; tail -3 gpurun_out/r2o_iter.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2o_iter_launches.csv python tools/mf_iter_bench.py --steps 2 > gpurun_out/r2o_ncu.log 2>&1
