#!/bin/bash
# ab.sh NAME...: for the in-tree library ("base") and each tools/variants/libpimdk_NAME.so: the bit-exactness tests of the
# CCpol gradient against the oracle, then per-kernel durations of one 32768-bead gradient pass (ncu, serialised)
mkdir -p gpurun_out
for n in base "$@"; do
  if [ $n = base ]; then unset PIMDK_LIB; else export PIMDK_LIB=$PWD/tools/variants/libpimdk_$n.so; fi
  if [ $n != base ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ccpol_energy_gradient_bit_exact or ccpol_all_surfaces" > gpurun_out/ab_$n.test 2>&1
    echo "$n tests: $(tail -1 gpurun_out/ab_$n.test)"
  fi
  bash tools/lv.sh $n $PIMDK_LIB | grep -E "sapt|sweep|rigid|total|saptx|saptsum"
done
