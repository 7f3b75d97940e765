#!/bin/bash
# C5 batch-scaling sweep (BASELINE configs[4]): CCpol-8sf, 1024 beads, 1k..32k trajectories on one GPU.
# usage: tools/c5_sweep.sh [out.jsonl]   (64k trajectories = 19 GB of state also fits; it takes ~2.5 min more)
OUT=${1:-gpurun_out/c5_sweep.jsonl}
: > $OUT
for nt in 1024 2048 4096 8192 16384 32768 ${C5_MAX:+65536}; do
  python bench.py --config c5 --ntraj $nt --steps 2 --warmup 3 --quick >> $OUT 2>> ${OUT%.jsonl}.err
done
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    if not l.startswith('{'): continue
    d = json.loads(l); print('%6d traj  %8.1f ms/step  %.3e bead-steps/s' % (d['config']['trajectories_per_gpu'], d['ms_per_step'], d['value']))"
