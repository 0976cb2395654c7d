#!/bin/bash
# experiment: k_mf_bounce_bwd capped at 128 / 102 registers (4 / 5 CTAs per SM instead of 3)
set -x
mkdir -p gpurun_out
for MB in 4 5; do
NMF_NVCC_EXTRA="-DNMF_BB_MINBLOCKS=$MB" timeout 900 python -m nmf_b200.build --force > gpurun_out/r2ao_build_$MB.log 2>&1; grep -A2 "k_mf_bounce_bwdILi0" gpurun_out/r2ao_build_$MB.log | grep -o "[0-9]* bytes spill stores\|Used [0-9]* registers" | paste - -
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2ao_mf_train_$MB.json 2>&1; tail -c 120 gpurun_out/r2ao_mf_train_$MB.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_mf_bounce' -c 12 --csv --log-file gpurun_out/r2ao_launches_$MB.csv python tools/mf_train_bench.py --steps 1 --retrace 1000 > /dev/null 2>&1
grep "k_mf_bounce" gpurun_out/r2ao_launches_$MB.csv | tail -2 | cut -d'"' -f10,28-32
done
