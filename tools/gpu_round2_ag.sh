#!/bin/bash
# Round 2, call AG: nmf_pack_shading (suite + training iteration breakdown) and the stand-alone device check log.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2ag_pytest_gpu.log; tail -4 gpurun_out/r2ag_pytest_gpu.log
timeout 300 python tools/mf_iter_bench.py --steps 20 > gpurun_out/r2ag_iter.json 2> gpurun_out/r2ag_iter.err; cat gpurun_out/r2ag_iter.json; tail -3 gpurun_out/r2ag_iter.err
timeout 120 tests/hostcheck/devcheck > gpurun_out/r2ag_devcheck.log 2>&1; tail -3 gpurun_out/r2ag_devcheck.log
