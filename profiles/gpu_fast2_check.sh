#!/bin/bash
# One GPU call: parity of the fast2 kernel, its timing next to the first-generation kernel, one ncu capture.
set -x
OUT=gpurun_out/r2
mkdir -p $OUT
python -m pytest tests/test_gpu_env_parity.py -x -q -m gpu -k "fast2 or full_size or bench_batch or large_shape_properties" > $OUT/pytest_fast2.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_fast2.log
tail -5 $OUT/pytest_fast2.log
python sweep.py --ns 100 --Bs 4096,16384 --rounds 3 > $OUT/sweep_v2.log 2>&1
python sweep.py --kernel-path 4 --ns 100 --Bs 4096 --rounds 3 > $OUT/sweep_v1.log 2>&1
grep env_kernel_ms $OUT/sweep_v2.log | cut -c1-200
grep env_kernel_ms $OUT/sweep_v1.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sap_real_fast2 --launch-skip 5 -c 1 -f -o $OUT/prof_fast2 python sweep.py --ns 100 --Bs 4096 --rounds 1 > $OUT/ncu_fast2.log 2>&1
tail -3 $OUT/ncu_fast2.log
