#!/bin/bash
# final tree of the third session: full GPU suite, ncu --set full of the strict pass (counters), bench lines (strict, fast, C5)
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2n_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ccpol_ -s 7 -c 7 -f -o $O/r2n_ccpol_strict python tools/prof_ccpol.py 0 32768 > $O/r2n_ncu_strict.log 2>&1; echo "ncu strict rc=$?"
bash tools/lv.sh final > /dev/null 2>&1
timeout 600 python bench.py > $O/r2n_bench_c4_strict.json 2> $O/r2n_bench_c4_strict.err; echo "c4 strict rc=$?"
timeout 600 python bench.py --mode fast --steps 3 --warmup 3 --no-cpu > $O/r2n_bench_c4_fast.json 2> $O/r2n_bench_c4_fast.err; echo "c4 fast rc=$?"
timeout 600 python bench.py --config c5 --steps 2 --warmup 3 > $O/r2n_bench_c5.json 2> $O/r2n_bench_c5.err; echo "c5 rc=$?"
python - <<PY
import json
for f in ('c4_strict','c4_fast','c5'):
    d=json.loads(open('$O/r2n_bench_%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'))
PY
