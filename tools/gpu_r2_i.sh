#!/bin/bash
# split SAPT pair sum: parity of the in-tree library, then per-kernel pass timings of the CTA-shape variants given as arguments
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ccpol" > $O/r2i_tests.log 2>&1; echo "ccpol tests rc=$?"; tail -3 $O/r2i_tests.log
bash tools/ab.sh "$@" 2>&1 | tee $O/r2i_ab.log
