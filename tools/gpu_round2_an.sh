#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_plugins.py tests/test_gpu_train.py -m gpu -q -k "device_resident or overflowed or fits_a_teacher" 2>&1 | tail -15 | cut -c1-400
