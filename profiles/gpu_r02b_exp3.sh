#!/bin/bash
# parity of the bench kernel after a change + A/B timing + phase timeline
OUT=gpurun_out/r2b
mkdir -p $OUT
export PYTHONPATH=$PWD
python -m pytest tests/test_gpu_env_parity.py tests/test_gpu_runner.py -x -q -m gpu -k "fast2 or full_size or bench_batch or fp16 or split or ahead or overlap or rollout_step or graph" > $OUT/pytest_fast2.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_fast2.log
tail -4 $OUT/pytest_fast2.log
bash profiles/gpu_ab.sh "" "--agent-in f16" "--agent-in none" > $OUT/ab.log 2>&1
BS=148,444 bash profiles/gpu_ab.sh "" >> $OUT/ab.log 2>&1
cat $OUT/ab.log
for B in 148 4096; do
  SAP_ABLATE=1 SAP_DEBUG_SKIP_REDO=99 python profiles/phase_timeline.py $B 2>&1 | head -10
done > $OUT/timeline.log 2>&1
cat $OUT/timeline.log
