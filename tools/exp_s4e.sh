#!/bin/bash
OUT=gpurun_out/s4e; mkdir -p $OUT
echo "== parity with the split kernel forced on"; GOF_FWD_SPLIT_MAX_TILES=1000000 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -5
for sp in 0 256 ; do
  echo "== single frame split_max=$sp"; GOF_FWD_SPLIT_MAX_TILES=$sp timeout 120 python tools/single_frame.py 256 100 2>&1 | tail -9 | cut -c1-170
done
echo "== 512^2 single frame"; for sp in 0 1024; do GOF_FWD_SPLIT_MAX_TILES=$sp timeout 120 python tools/single_frame.py 512 50 2>&1 | tail -1; done
echo "== batched with split forced"; GOF_FWD_SPLIT_MAX_TILES=1000000 bash tools/ab.sh $OUT splitall=
GOF_FWD_SPLIT_MAX_TILES=256 bash tools/ab.sh $OUT split256=
