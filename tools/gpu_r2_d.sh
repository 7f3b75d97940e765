#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/r2d_tests.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:agrad_ -s 4 -c 4 -f -o $O/r2d_ccpol_analytic python tools/prof_ccpol.py 2 262144 > $O/r2d_ncu_analytic.log 2>&1; echo "ncu analytic rc=$?"
timeout 600 python bench.py --mode analytic --steps 5 --warmup 3 > $O/r2d_bench_c4_analytic.json 2> $O/r2d_bench_c4_analytic.err; echo "c4 analytic rc=$?"
head -c 300 $O/r2d_bench_c4_analytic.json; echo
