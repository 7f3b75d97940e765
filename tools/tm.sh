#!/bin/bash
# tm.sh NAME...: multi-pass strict gradient timing (tools/time_ccpol.py) of the in-tree library and the named variants
mkdir -p gpurun_out
for n in base "$@"; do
  if [ $n = base ]; then unset PIMDK_LIB; else export PIMDK_LIB=$PWD/tools/variants/libpimdk_$n.so; fi
  python tools/time_ccpol.py 0 262144 2>&1 | tail -1
done | tee -a gpurun_out/tm.log
