#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / transient): ./profiles/gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
exit $rc
