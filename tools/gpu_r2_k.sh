#!/bin/bash
# tables as kernel parameters: A/B against the HEAD library on one box, full GPU suite, default bench
mkdir -p gpurun_out; O=gpurun_out
bash tools/ab.sh head 2>&1 | tee $O/r2k_ab.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2k_tests.log
timeout 600 python bench.py > $O/r2k_bench_default.json 2> $O/r2k_bench_default.err; echo "default rc=$?"
python - <<PY
import json
d=json.loads(open('$O/r2k_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'))
PY
