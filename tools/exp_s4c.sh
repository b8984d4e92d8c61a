#!/bin/bash
OUT=gpurun_out/s4c; mkdir -p $OUT
V=/root/repo/f3d_gaus_b200/variants
echo "== parity with the wide kernel forced on"; GOF_FWD_WIDE_MAX_TILES=100000 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3
echo "== batched A/B"; bash tools/ab.sh $OUT base=$V/libgof_b200_base.so new= sw4=$V/libgof_b200_sw4.so
for spec in base:$V/libgof_b200_base.so:0 new:$V/../libgof_b200.so:0 new:$V/../libgof_b200.so:300 sw4:$V/libgof_b200_sw4.so:0 sw4:$V/libgof_b200_sw4.so:300 np3:$V/libgof_b200_np3.so:300; do
  IFS=: read label lib wide <<< "$spec"
  echo "== single frame $label wide=$wide"; GOF_B200_LIB=$lib GOF_FWD_WIDE_MAX_TILES=$wide timeout 120 python tools/single_frame.py 256 100 2>&1 | tail -9 | cut -c1-170
done
