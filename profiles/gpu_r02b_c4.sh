#!/bin/bash
# C4 shape (64 x 324 x 450) and other multi-CTA shapes: parity + A/B timing (compiled-in shape vs run-time shape)
OUT=gpurun_out/r2b
mkdir -p $OUT
export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_env_parity.py -x -q -m gpu -k "constellation or multi_cta or full_size or large" 2>&1 | tail -3
for path in 0 5; do python profiles/time_env_step.py 64 324 450 $path; python profiles/time_env_step.py 256 324 450 $path; done 2>&1 | tee $OUT/ab_c4.log
