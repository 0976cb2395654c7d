#!/bin/bash
# Round 2, call AU: source-level ncu capture of k_march<0> (instruction mix of pass A / pass B).
set -x
mkdir -p gpurun_out /tmp/prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_march' -c 1 -o /tmp/prof/march python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2au_ncu.log 2>&1
ncu -i /tmp/prof/march.ncu-rep --page source --csv > gpurun_out/r2au_march_src.csv 2>/dev/null
ncu -i /tmp/prof/march.ncu-rep --page raw --csv > gpurun_out/r2au_march_raw.csv 2>/dev/null
ls -la gpurun_out | grep r2au
