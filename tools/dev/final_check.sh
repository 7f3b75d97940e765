S=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/v23_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s)-S )) s" > gpurun_out/v23_times.txt
timeout 240 python bench.py > gpurun_out/v23_bench.json 2> gpurun_out/v23_bench.err; echo "bench rc=$? $(( $(date +%s)-S )) s" >> gpurun_out/v23_times.txt
timeout 200 compute-sanitizer --tool memcheck python tools/dev/check_streams.py 34 > gpurun_out/v23_memcheck_streams.log 2>&1; echo "memcheck rc=$? $(( $(date +%s)-S )) s" >> gpurun_out/v23_times.txt
timeout 150 compute-sanitizer --tool memcheck python tools/dev/sanitize.py > gpurun_out/v23_memcheck.log 2>&1; echo "memcheck2 rc=$? $(( $(date +%s)-S )) s" >> gpurun_out/v23_times.txt
tail -2 gpurun_out/v23_tests.log; cat gpurun_out/v23_times.txt; tail -3 gpurun_out/v23_memcheck_streams.log; tail -3 gpurun_out/v23_memcheck.log
timeout 170 ncu --set full --clock-control none --import-source on -k regex:ccpol_ -s 7 -c 7 -f -o gpurun_out/ccpol_pipeline_v23 python tools/dev/prof_ccpol.py 0 32768 > gpurun_out/v23_ncu.log 2>&1; echo "ncu rc=$? $(( $(date +%s)-S )) s" >> gpurun_out/v23_times.txt; tail -1 gpurun_out/v23_times.txt; ls -la gpurun_out/*.ncu-rep
