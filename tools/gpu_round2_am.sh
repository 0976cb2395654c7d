#!/bin/bash
# Round 2, call AM: device-resident environment scalars (NmfScene.env_dyn): suite + training iteration.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2am_pytest_gpu.log; tail -25 gpurun_out/r2am_pytest_gpu.log | cut -c1-300
timeout 300 python tools/mf_iter_bench.py --steps 20 > gpurun_out/r2am_iter.json 2> gpurun_out/r2am_iter.err; cat gpurun_out/r2am_iter.json; tail -3 gpurun_out/r2am_iter.err
timeout 300 python bench.py --no-cpu --no-train --no-refcuda --sustain-s 0 --steps 5 --warmup 3 > gpurun_out/r2am_bench.json 2> gpurun_out/r2am_bench.err; python tools/bench_phases.py gpurun_out/r2am_bench.json
