#!/bin/bash
# full GPU validation: test suite, smoke, default bench line (what the driver runs at round end)
export OUT=${OUT:-gpurun_out/r2}
mkdir -p $OUT
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 400 $OUT/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref2.json 2>$OUT/bench_ref2.err; cut -c1-400 $OUT/bench_ref2.json
python - <<'PY'
import json
import os
d=json.loads(open(os.environ.get("OUT", "gpurun_out/r2") + "/bench_default.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["wall_s"])
PY
