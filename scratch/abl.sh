#!/bin/bash
# timing ablations of the real-env fast kernel: SAP_DEBUG_SKIP_REDO bits (see sap_real_fast.cu)
for f in 0 256 512 128 24 ; do
  echo -n "flags=$f  "; SAP_DEBUG_SKIP_REDO=$f bash scratch/kbench.sh ${1:-c3}
done
