#!/bin/bash
# C2 shape (1024 x 50 x 50): parity of the n <= 64 kernel + A/B timing (compiled-in shape vs run-time shape)
OUT=gpurun_out/r2b
mkdir -p $OUT
export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_env_parity.py -x -q -m gpu -k "golden or oracle" 2>&1 | tail -3
NS=50 BS=1024,4096 bash profiles/gpu_ab.sh "" "--kernel-path 5" 2>&1 | tee $OUT/ab_c2.log
