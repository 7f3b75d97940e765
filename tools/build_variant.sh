#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..."  -> tools/variants/libpimdk_NAME.so (strict CCpol object rebuilt with the flags;
# every other object taken from the in-tree build)
set -e
R=$(cd $(dirname $0)/.. && pwd)
mkdir -p $R/tools/variants
O=/tmp/variant_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off -fmad=false -DPIMDK_CCPOL_STRICT=1 $2 -c $R/pimd_tunneling_b200/csrc/ccpol_kernels.cu -o $O
B=$R/pimd_tunneling_b200/build
OBJS=$(ls $B/*.o | grep -v ccpol_strict.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $R/tools/variants/libpimdk_$1.so $O $OBJS $EXTRA_OBJS -cudart=shared -ldl
echo built $1
