#!/bin/bash
# A/B timing of the env kernel under environment-variable switches: gpu_ab.sh "VAR=val VAR2=val" ...
OUT=gpurun_out/r2
mkdir -p $OUT
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python sweep.py --ns 100 --Bs 4096 --rounds 5 --env-only 2>&1 | grep '"env_kernel_ms"' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['B_run'], r['n'], 'ms', r['env_kernel_ms'], 'frac', round(r['hbm_frac'],3))"
done
