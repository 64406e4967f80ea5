#!/bin/bash
for f in ${FLAGS:-0 1024 256 1280}; do
  echo -n "flags=$f  "; SAP_DEBUG_SKIP_REDO=$f bash scratch/kbench.sh ${1:-c3}
done
