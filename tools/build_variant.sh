#!/bin/bash
# build a kernel-variant library for A/B experiments: tools/build_variant.sh <name> <extra nvcc flags...>
set -e
NAME=$1; shift
D=/root/repo/f3d_gaus_b200/csrc
OUT=/root/repo/f3d_gaus_b200/variants; mkdir -p $OUT/obj_$NAME
for f in abi preprocess binning render_fwd render_bwd preprocess_bwd integrate predictor_head epilogue; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $D/$f.cu -o $OUT/obj_$NAME/$f.o &
done; wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libgof_b200_$NAME.so $OUT/obj_$NAME/*.o -lcudart
echo built $OUT/libgof_b200_$NAME.so
