#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_mf_train.py -m gpu -q -k "edge" 2>&1 | tail -15
