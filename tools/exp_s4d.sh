#!/bin/bash
OUT=gpurun_out/s4d; mkdir -p $OUT
V=/root/repo/f3d_gaus_b200/variants
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -3
echo "== batched A/B"; bash tools/ab.sh $OUT base=$V/libgof_b200_base.so new= sw3=$V/libgof_b200_sw3.so sw4=$V/libgof_b200_sw4.so sw6=$V/libgof_b200_sw6.so
for spec in base:$V/libgof_b200_base.so new:$V/../libgof_b200.so sw4:$V/libgof_b200_sw4.so; do
  IFS=: read label lib <<< "$spec"
  echo "== single frame $label"; GOF_B200_LIB=$lib timeout 120 python tools/single_frame.py 256 100 2>&1 | tail -9 | cut -c1-170
done
python tools/per_view_timeline.py 256 2>&1 | tail -4
