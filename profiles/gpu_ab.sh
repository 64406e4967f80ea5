#!/bin/bash
# A/B timing of the env kernel: every argument is a list of extra sweep.py options, e.g.
#   gpu_ab.sh "" "--kernel-path 4" "--agent-in f16" "--agent-in none"
# (prefix an argument with VAR=value words to set environment variables for that run)
for cfg in "$@"; do
  echo "== ${cfg:-default}"
  envs=""; args=""
  for w in $cfg; do case "$w" in [A-Z_]*=*) envs="$envs $w";; *) args="$args $w";; esac; done
  env $envs python sweep.py --ns ${NS:-100} --Bs ${BS:-4096} --rounds 5 --env-only $args 2>&1 | grep '"env_kernel_ms"' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['B_run'], r['n'], 'ms', r['env_kernel_ms'], 'frac', round(r['hbm_frac'],3), 'frac_8d', round(r['hbm_frac_8d'],3), 'agent_in', r['agent_in'])"
done
