#!/bin/bash
# ab_noparity.sh NAME...: per-kernel durations only (timing experiments whose results are not meant to be right)
mkdir -p gpurun_out
for n in "$@"; do
  export PIMDK_LIB=$PWD/tools/variants/libpimdk_$n.so
  bash tools/lv.sh $n $PIMDK_LIB | grep -E "${AB_FILTER:-sapt|sweep|rigid|total}"
done
