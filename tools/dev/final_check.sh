S=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/v22_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s)-S )) s" > gpurun_out/v22_times.txt
timeout 240 python bench.py > gpurun_out/v22_bench.json 2> gpurun_out/v22_bench.err; echo "bench rc=$? $(( $(date +%s)-S )) s" >> gpurun_out/v22_times.txt
timeout 200 compute-sanitizer --tool memcheck python tools/dev/check_streams.py 34 > gpurun_out/v22_memcheck_streams.log 2>&1; echo "memcheck rc=$? $(( $(date +%s)-S )) s" >> gpurun_out/v22_times.txt
timeout 200 compute-sanitizer --tool racecheck python tools/dev/check_streams.py 34 > gpurun_out/v22_racecheck_streams.log 2>&1; echo "racecheck rc=$? $(( $(date +%s)-S )) s" >> gpurun_out/v22_times.txt
tail -2 gpurun_out/v22_tests.log; cat gpurun_out/v22_times.txt; tail -3 gpurun_out/v22_memcheck_streams.log; tail -3 gpurun_out/v22_racecheck_streams.log
