#!/bin/bash
# Round 2, call F: full GPU suite (new: forest env, reference-written checkpoint, estimator distribution, mipbias at full
# size, occupancy with existing mask) with the parity reports printed; gather ceiling by segment size.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/r2f_pytest_gpu_full.log; tail -15 gpurun_out/r2f_pytest_gpu_full.log
timeout 120 python -c "
from nmf_b200 import ops
import json
print(json.dumps({f'{16*g}B_l2_78MB': ops.gather_peak(78<<20, group=g) for g in (1,2,4,8)} | {f'{16*g}B_dram_8GB': ops.gather_peak(8<<30, taps=32, group=g) for g in (1,4,8)}))
" > gpurun_out/r2f_gather_peaks.json 2>&1; cat gpurun_out/r2f_gather_peaks.json
