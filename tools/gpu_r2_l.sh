#!/bin/bash
# ncu --set full of the seven kernels of one strict gradient pass (second pass of the run)
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-r2l}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ccpol_ -s 7 -c 7 -f -o $O/${TAG}_ccpol_strict python tools/prof_ccpol.py 0 32768 > $O/${TAG}_ncu_strict.log 2>&1; echo "ncu strict rc=$?"
ls -la $O | grep ${TAG}_
