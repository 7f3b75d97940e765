#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/r2f_tests.log
timeout 300 python bench.py --config c2 --steps 10000 --warmup 500 > $O/r2f_bench_c2.json 2> $O/r2f_bench_c2.err; echo "c2 rc=$?"
timeout 300 python tools/prof_path.py --pes 2dtest --n 256 --ntraj 4096 --thermostat 1 --steps 200 --noutput 100 > $O/r2f_fam_c2.txt 2>&1; cat $O/r2f_fam_c2.txt
timeout 600 python bench.py > $O/r2f_bench_default.json 2> $O/r2f_bench_default.err; echo "default rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_reference.json 2> $O/r2f_bench_reference.err; echo "reference rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2f_smoke.log
python -c "
import json
for f in ('c2','default','reference'):
    d=json.loads(open('$O/r2f_bench_%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('cpu_baseline'))
"
