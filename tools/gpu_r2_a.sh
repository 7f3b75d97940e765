#!/bin/bash
# round 2, GPU call A: parity suite, bench lines of C4/C2/C1, family timings, ncu --set full of the HBM-shaped kernels
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2a_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r2a_bench_c4.json 2> $O/r2a_bench_c4.err; echo "c4 rc=$?"
timeout 300 python bench.py --config c2 --steps 200 --warmup 20 > $O/r2a_bench_c2.json 2> $O/r2a_bench_c2.err; echo "c2 rc=$?"
timeout 300 python bench.py --config c1 --steps 2000 --warmup 200 > $O/r2a_bench_c1.json 2> $O/r2a_bench_c1.err; echo "c1 rc=$?"
timeout 300 python tools/prof_path.py --ntraj 2048 --steps 2 > $O/r2a_fam_c4shape.txt 2>&1
timeout 300 python tools/prof_path.py --pes 2dtest --n 256 --ntraj 4096 --thermostat 1 --steps 100 --noutput 100 > $O/r2a_fam_c2.txt 2>&1
timeout 300 python tools/prof_path.py --pes 2dtest --n 256 --ntraj 4096 --thermostat 1 --steps 100 --noutput 100 --gemm 3 > $O/r2a_fam_c2_g3.txt 2>&1
timeout 300 python tools/prof_path.py --n 1024 --ntraj 1024 --steps 1 > $O/r2a_fam_c5shape.txt 2>&1
cat $O/r2a_fam_*.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:'nm_update2|estimator_kernel|nm_gemm_pipe|beadvec' -c 10 -f -o $O/r2a_nm_pile python tools/prof_path.py --ntraj 2048 --steps 2 --warm 0 > $O/r2a_ncu_pile.log 2>&1; echo "ncu pile rc=$?"
timeout 600 $NCU -k regex:'sample_momenta2|estimator_modes|andersen_clock|nm_update2|add_kernel' -c 12 -f -o $O/r2a_nm_andersen python tools/prof_path.py --ntraj 2048 --thermostat 1 --noutput 1 --steps 2 --warm 0 > $O/r2a_ncu_andersen.log 2>&1; echo "ncu andersen rc=$?"
timeout 600 $NCU -k regex:'ccpol_sites|ccpol_combine|ccpol_setup|ccpol_dipind' -s 8 -c 4 -f -o $O/r2a_pes_hbm python tools/prof_ccpol.py 0 32768 > $O/r2a_ncu_peshbm.log 2>&1; echo "ncu pes rc=$?"
ls -la $O | grep r2a
