#!/bin/bash
# usage: tools/gpu_r2_multi.sh N   — multi-GPU measurements on N GPUs of one box (N = 1: the sweep's single-GPU leg only)
N=$1; mkdir -p gpurun_out; O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"
[ $N -eq 1 ] && RUN="python"
if [ $N -ge 2 ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -x -q > $O/r2m_test_${N}gpu.log 2>&1; echo "mgpu test rc=$?"; tail -3 $O/r2m_test_${N}gpu.log
  timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > $O/r2m_c4_weak_${N}gpu.json 2> $O/r2m_c4_weak_${N}gpu.err; echo "c4 weak rc=$?"
  timeout 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --scaling strong > $O/r2m_c4_strong_${N}gpu.json 2> $O/r2m_c4_strong_${N}gpu.err; echo "c4 strong rc=$?"
fi
timeout 2400 $RUN bench.py --gpus $N --config c5 --scaling strong --sweep 1024,2048,4096,8192,16384,32768,65536 --steps 2 --warmup 1 > $O/r2m_c5_sweep_${N}gpu.jsonl 2> $O/r2m_c5_sweep_${N}gpu.err; echo "c5 sweep rc=$?"
python - <<PY
import json
for f in ("$O/r2m_c4_weak_${N}gpu.json", "$O/r2m_c4_strong_${N}gpu.json", "$O/r2m_c5_sweep_${N}gpu.jsonl"):
    try:
        for l in open(f):
            if l.startswith("{"):
                d = json.loads(l); c = d["config"]
                print(f.split("/")[-1], d["n_gpus"], d.get("scaling"), c.get("trajectories_total"), c.get("trajectories_per_gpu", c.get("trajectories_this_gpu")), "%.1f ms/step" % d["ms_per_step"], "%.4g bead-steps/s" % d["value"])
    except Exception as e:
        print(f, e)
PY
tail -3 $O/r2m_*_${N}gpu.err | tail -20
