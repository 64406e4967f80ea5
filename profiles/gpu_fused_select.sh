#!/bin/bash
# fused selection + step (sap_rollout_step): tests, then the bench with its variants
OUT=gpurun_out/r2
mkdir -p $OUT
python -m pytest tests/test_gpu_runner.py -x -q -m gpu -k "fused_select or rollout_step_entry or overlapped or graph or oracle" 2>&1 | tail -15
python bench.py --no-cpu > $OUT/bench_fsel.json 2> $OUT/bench_fsel.err; tail -c 600 $OUT/bench_fsel.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_fsel.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["impl_config"]["fuse_select_step"], d["e2e"]["value"] if d["e2e"] else None)
for k,v in (d.get("variants") or {}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("env_kernel_ms"), v.get("error"))
PY
