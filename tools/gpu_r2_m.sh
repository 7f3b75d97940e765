#!/bin/bash
# analytic mode after the kernel-parameter tables: tests, timing, ncu capture for the counters, bench line
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "analytic" > $O/r2m_tests.log 2>&1; echo "analytic tests rc=$?"; tail -2 $O/r2m_tests.log
python tools/time_ccpol.py 2 1048576 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:agrad_ -s 4 -c 4 -f -o $O/r2m_ccpol_analytic python tools/prof_ccpol.py 2 262144 > $O/r2m_ncu_analytic.log 2>&1; echo "ncu analytic rc=$?"
timeout 600 python bench.py --mode analytic --steps 5 --warmup 3 > $O/r2m_bench_c4_analytic.json 2> $O/r2m_bench_c4_analytic.err; echo "c4 analytic rc=$?"
head -c 400 $O/r2m_bench_c4_analytic.json; echo
