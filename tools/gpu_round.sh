#!/bin/bash
# One GPU session: bench line, ncu launch list, full ncu capture of the main kernels.  Run under gpurun.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${REF:-0}" = "1" ]; then
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
grep -c k_ gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march|k_shade|k_bounce|k_incoming|k_select|k_reduce0' -c 11 \
    -o gpurun_out/prof python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
fi
