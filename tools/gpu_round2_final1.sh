#!/bin/bash
# Round 2, final 1-GPU measurement: suite, full bench line, reference arms, training step + iteration, ncu launch lists and
# one --set full capture of the render kernels and of one training step (exported to CSV on the box).
set -x
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2f_pytest_gpu.log; tail -4 gpurun_out/r2f_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python tools/bench_phases.py gpurun_out/r2f_bench.json; tail -2 gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_ref.err; tail -c 400 gpurun_out/r2f_bench_reference_arm.json
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 1000 > gpurun_out/r2f_mf_train.json 2>&1; tail -c 300 gpurun_out/r2f_mf_train.json
timeout 300 python tools/mf_train_bench.py --steps 20 --retrace 38000 > gpurun_out/r2f_mf_train_38k.json 2>&1; tail -c 300 gpurun_out/r2f_mf_train_38k.json
timeout 300 python tools/mf_iter_bench.py --steps 20 > gpurun_out/r2f_iter.json 2> gpurun_out/r2f_iter.err; cat gpurun_out/r2f_iter.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2f_ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 200 --csv --log-file gpurun_out/r2f_train_launches.csv \
    python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2f_ncu_train.log 2>&1
EXTRA=lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum
timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_march|k_shade|k_bounce|k_incoming|k_select|k_reduce0' -c 10 \
    -o /tmp/prof/render python bench.py --steps 1 --warmup 1 --no-cpu --no-train --no-refcuda --sustain-s 0 > gpurun_out/r2f_ncu_render.log 2>&1
ncu -i /tmp/prof/render.ncu-rep --page raw --csv > gpurun_out/r2f_render_raw.csv 2>/dev/null
timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_mf_|k_select|k_train_sample' -c 16 \
    -o /tmp/prof/train python tools/mf_train_bench.py --steps 1 --retrace 1000 > gpurun_out/r2f_ncu_trainfull.log 2>&1
ncu -i /tmp/prof/train.ncu-rep --page raw --csv > gpurun_out/r2f_train_raw.csv 2>/dev/null
ls -la gpurun_out | grep r2f; du -sh gpurun_out
