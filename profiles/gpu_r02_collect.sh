#!/bin/bash
# Round-2 evidence in one GPU call: bench lines, ncu launch list of the bench command, one full ncu capture of the env
# kernel, the C5 sweep.  Every profiled command first runs to completion without ncu.
OUT=${OUT:-gpurun_out/r2}
mkdir -p $OUT
run() { name=$1; shift; python bench.py "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err; grep -c '"metric"' $OUT/bench_$name.json; }
run c3 ; run c2 --workload c2 --no-variants; run c4 --workload c4 --no-variants; run c1 --workload c1 --no-variants; run c3mock --workload c3mock --no-variants
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-variants > $OUT/plain.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-variants --no-graph > $OUT/ncu_launches.log 2>&1
python sweep.py --ns 100 --Bs 4096 --rounds 1 --env-only > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sap_real_fast2 --launch-skip 5 -c 1 -f -o $OUT/prof_fast2_final python sweep.py --ns 100 --Bs 4096 --rounds 1 --env-only > $OUT/ncu_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sap_real_fast2 --launch-skip 5 -c 1 -f -o $OUT/prof_fast2_final_f16 python sweep.py --ns 100 --Bs 4096 --rounds 1 --env-only --agent-in f16 > $OUT/ncu_final16.log 2>&1
timeout 900 python sweep.py --cpu --out $OUT/r02_sweep_c5.jsonl > $OUT/sweep.log 2>&1
tail -30 $OUT/sweep.log
