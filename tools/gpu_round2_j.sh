#!/bin/bash
# Round 2, call J: suite (all-BF16 reverse MLP), then ncu --set full captures: render kernels of one bench image, kernels of
# one microfacet training step; launch list of the render.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2j_pytest_gpu.log; tail -6 gpurun_out/r2j_pytest_gpu.log
EXTRA=lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum
timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_march|k_shade|k_bounce|k_incoming|k_select|k_reduce0' -c 10 \
    -o gpurun_out/r2j_prof_render python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2j_ncu_render.log 2>&1
timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_mf_|k_train_sample|k_train_prefix' -c 16 \
    -o gpurun_out/r2j_prof_train python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2j_ncu_train.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2j_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2j_ncu_launches.log 2>&1
ls -la gpurun_out | grep r2j
