#!/bin/bash
# quick kernel-time probe: prints value, ms/step, env-kernel ms
python bench.py --workload ${1:-c3} --steps 2 --warmup 3 --no-cpu --no-e2e ${2:-} 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  ms/step %.1f  kernel_ms %.4f  share %.3f' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['kernel_share_of_step']))"
