#!/bin/bash
# Round 2: ncu --set full of the LAST build (render kernels of one bench image, kernels of one microfacet training step).
set -x
mkdir -p gpurun_out /tmp/prof
EXTRA=lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum
timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_march|k_shade|k_bounce|k_incoming|k_select|k_reduce0' -c 10 \
    -o /tmp/prof/render python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2k4_ncu_render.log 2>&1
ncu -i /tmp/prof/render.ncu-rep --page raw --csv > gpurun_out/r2k4_render_raw.csv 2>/dev/null
timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_mf_|k_select' -c 16 \
    -o /tmp/prof/train python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2k4_ncu_train.log 2>&1
ncu -i /tmp/prof/train.ncu-rep --page raw --csv > gpurun_out/r2k4_train_raw.csv 2>/dev/null
ls -la gpurun_out | grep r2k2
