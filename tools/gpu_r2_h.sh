#!/bin/bash
# full verification of the tree: GPU tests, smoke, default bench + reference arm, memcheck of the sanitizer workload
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2h_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > $O/r2h_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/r2h_memcheck.log
timeout 600 python bench.py > $O/r2h_bench_default.json 2> $O/r2h_bench_default.err; echo "default rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2h_bench_reference.json 2> $O/r2h_bench_reference.err; echo "reference rc=$?"
python - <<PY
import json
for f in ('default','reference'):
    d=json.loads(open('$O/r2h_bench_%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'))
PY
