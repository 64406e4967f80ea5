#!/bin/bash
# phase timeline + full ncu capture of the shape-specialised bench kernel
OUT=gpurun_out/r2b
mkdir -p $OUT
export PYTHONPATH=$PWD
for B in 148 4096; do
  SAP_ABLATE=1 SAP_DEBUG_SKIP_REDO=99 python profiles/phase_timeline.py $B 2>&1 | head -12
done > $OUT/timeline_fixed.log 2>&1
cat $OUT/timeline_fixed.log
python sweep.py --ns 100 --Bs 4096 --rounds 1 --env-only > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sap_real_fast2 --launch-skip 5 -c 1 -f -o $OUT/prof_fast2_fixed python sweep.py --ns 100 --Bs 4096 --rounds 1 --env-only > $OUT/ncu_fixed.log 2>&1
tail -2 $OUT/ncu_fixed.log
