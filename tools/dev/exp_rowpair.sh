# row-pair experiment: bit-exactness of a variant library, then per-kernel timings (usage: exp_rowpair.sh NAME)
V=$PWD/tools/dev/variants/libpimdk_$1.so
PIMDK_LIB=$V timeout 200 python -m pytest tests -m gpu -x -q -k "ccpol or smoke or golden or bit" > gpurun_out/exp_$1_tests.log 2>&1; echo "variant tests rc=$?"; tail -1 gpurun_out/exp_$1_tests.log
bash tools/dev/lv.sh $1 $V 2>&1 | tail -9
