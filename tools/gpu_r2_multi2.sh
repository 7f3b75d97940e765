#!/bin/bash
# usage: tools/gpu_r2_multi2.sh N — multi-GPU sanity of the final tree: 2-rank NCCL test, C4 weak and strong bench lines
N=$1; mkdir -p gpurun_out; O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q > $O/r2o_test_${N}gpu.log 2>&1; echo "mgpu test rc=$?"; tail -2 $O/r2o_test_${N}gpu.log
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > $O/r2o_c4_weak_${N}gpu.json 2> $O/r2o_c4_weak_${N}gpu.err; echo "c4 weak rc=$?"
timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --scaling strong > $O/r2o_c4_strong_${N}gpu.json 2> $O/r2o_c4_strong_${N}gpu.err; echo "c4 strong rc=$?"
python - <<PY
import json
for f in ("$O/r2o_c4_weak_${N}gpu.json", "$O/r2o_c4_strong_${N}gpu.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); c = d["config"]
            print(f.split("/")[-1], d["n_gpus"], d.get("scaling"), c.get("trajectories_total"), "%.1f ms/step" % d["ms_per_step"], "%.4g bead-steps/s" % d["value"], "e2e %.4g" % d["e2e"]["value"])
PY
