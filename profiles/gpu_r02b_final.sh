#!/bin/bash
# final C3 evidence of the round: default bench line + full ncu capture of the bench kernel (after the same command ran without ncu)
OUT=gpurun_out/r2e
mkdir -p $OUT
export PYTHONPATH=$PWD
python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err; grep -c '"metric"' $OUT/bench_c3.json
python sweep.py --ns 100 --Bs 4096 --rounds 1 --env-only > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sap_real_fast2 --launch-skip 5 -c 1 -f -o $OUT/prof_fast2_final python sweep.py --ns 100 --Bs 4096 --rounds 1 --env-only > $OUT/ncu_final.log 2>&1
SAP_ABLATE=1 SAP_DEBUG_SKIP_REDO=99 python profiles/phase_timeline.py 4096 2>&1 | head -12 > $OUT/timeline.log
SAP_ABLATE=1 SAP_DEBUG_SKIP_REDO=99 python profiles/phase_timeline.py 148 2>&1 | head -10 >> $OUT/timeline.log
cat $OUT/timeline.log
