#!/bin/bash
# ncu --set full of the two kernels of the split pair sum (one launch each, second pass)
mkdir -p gpurun_out; O=gpurun_out
[ -n "$1" ] && export PIMDK_LIB=$PWD/tools/variants/libpimdk_$1.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sapt -s 2 -c 2 -f -o $O/r2j_sapt_split${1:+_$1} python tools/prof_ccpol.py 0 32768 > $O/r2j_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O | grep r2j
