#!/bin/bash
# round 2, GPU call B: parity suite incl. the analytic-gradient mode, rigid-stage A/B, ncu --set full of the strict and analytic
# pipelines, bench lines (analytic, fast, C2, C1, C5), ncu launch list of the default bench
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2b_tests.log
for v in base disprow disprow4; do
  if [ $v = base ]; then bash tools/lv.sh $v 2>&1 | tail -9; else bash tools/lv.sh $v $PWD/tools/variants/libpimdk_$v.so 2>&1 | grep -E "rigid|total"; fi
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:ccpol_ -s 7 -c 7 -f -o $O/r2b_ccpol_strict python tools/prof_ccpol.py 0 32768 > $O/r2b_ncu_strict.log 2>&1; echo "ncu strict rc=$?"
timeout 600 $NCU -k regex:agrad_ -s 4 -c 4 -f -o $O/r2b_ccpol_analytic python tools/prof_ccpol.py 2 262144 > $O/r2b_ncu_analytic.log 2>&1; echo "ncu analytic rc=$?"
timeout 600 python bench.py --mode analytic --steps 5 --warmup 3 > $O/r2b_bench_c4_analytic.json 2> $O/r2b_bench_c4_analytic.err; echo "c4 analytic rc=$?"
timeout 600 python bench.py --mode fast --steps 3 --warmup 3 --no-cpu > $O/r2b_bench_c4_fast.json 2> $O/r2b_bench_c4_fast.err; echo "c4 fast rc=$?"
timeout 600 python bench.py --config c5 --steps 2 --warmup 3 > $O/r2b_bench_c5.json 2> $O/r2b_bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --config c2 --steps 10000 --warmup 500 > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --config c1 --steps 500000 --warmup 20000 > $O/r2b_bench_c1.json 2> $O/r2b_bench_c1.err; echo "c1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2b_launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu --ntraj 2048 > $O/r2b_launches_c4.log 2>&1; echo "launch list rc=$?"
head -c 600 $O/r2b_bench_c4_analytic.json; echo
ls -la $O | grep r2b
