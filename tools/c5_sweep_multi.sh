#!/bin/bash
# C5 batch-scaling sweep at N GPUs of one box (weak scaling: trajectories PER GPU fixed as N grows).
# usage: tools/c5_sweep_multi.sh N [out.jsonl] [max trajectories per GPU, default 8192]
N=$1; OUT=${2:-gpurun_out/c5_sweep_${N}gpu.jsonl}; MAX=${3:-8192}
: > $OUT
for nt in 1024 2048 4096 8192 16384 32768; do
  [ $nt -gt $MAX ] && break
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --config c5 --ntraj $nt --steps 2 --warmup 3 --quick >> $OUT 2>> ${OUT%.jsonl}.err
done
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d = json.loads(l); print('%d GPU x %6d traj  %8.1f ms/step  %.3e bead-steps/s' % (d['n_gpus'], d['config']['trajectories_per_gpu'], d['ms_per_step'], d['value']))"
