#!/bin/bash
# Round 2 (second session), experiment 1: shape-specialised instantiation of the bench kernel (n = m = 100 compiled in)
# next to the run-time-shape one, and the per-CTA phase timeline at 1, 2 and 3 resident CTAs per SM.
set -x
OUT=gpurun_out/r2b
mkdir -p $OUT
python -m pytest tests/test_gpu_env_parity.py -x -q -m gpu -k "fast2 or full_size or bench_batch" > $OUT/pytest_fast2.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_fast2.log
tail -3 $OUT/pytest_fast2.log
bash profiles/gpu_ab.sh "" "--kernel-path 5" "--agent-in f16" "--agent-in none" > $OUT/ab1.log 2>&1
BS=148,296,444,888 bash profiles/gpu_ab.sh "" "--kernel-path 5" >> $OUT/ab1.log 2>&1
cat $OUT/ab1.log
for B in 148 296 444; do
  SAP_ABLATE=1 SAP_DEBUG_SKIP_REDO=99 python profiles/phase_timeline.py $B 2>&1 | head -12
done > $OUT/timeline_small.log 2>&1
cat $OUT/timeline_small.log
