#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -12 $O/r2e_tests.log
timeout 300 python bench.py --config c2 --steps 10000 --warmup 500 > $O/r2e_bench_c2.json 2> $O/r2e_bench_c2.err; echo "c2 rc=$?"
timeout 300 python tools/prof_path.py --pes 2dtest --n 256 --ntraj 4096 --thermostat 1 --steps 200 --noutput 100 > $O/r2e_fam_c2.txt 2>&1; cat $O/r2e_fam_c2.txt
python -c "
import json
d=json.loads(open('$O/r2e_bench_c2.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other_kernels_ms'])
"
