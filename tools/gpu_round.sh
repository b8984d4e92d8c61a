#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, kernel launch list, ncu full captures of the pipeline kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 python bench.py > $OUT/bench_ours.json 2> $OUT/bench_ours.err
timeout 300 python tools/quick_bench.py 256 256 50 > $OUT/quick_bench.log 2>&1; cat $OUT/quick_bench.log
timeout 200 python tools/single_frame.py 256 100 > $OUT/single_frame.log 2>&1; tail -1 $OUT/single_frame.log
timeout 200 python tools/per_view_timeline.py 256 2>&1 | grep render_predicted > $OUT/per_view_timeline.log; cat $OUT/per_view_timeline.log
(timeout 120 python tools/bench_head.py 8 256; timeout 120 python tools/bench_head.py 64 256) > $OUT/head_bench.json 2> $OUT/head_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > $OUT/launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_train.csv \
    python bench.py --workload train256 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'render_fwd_kernel|tile_sort_gather|preprocess_kernel|scatter_kernel|tile_scan' -s 15 -c 5 \
    -o $OUT/prof_blend python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'render_bwd_kernel|preprocess_bwd' -s 8 -c 2 \
    -o $OUT/prof_bwd python bench.py --workload train256 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'render_fwd_split|tile_sort_gather|preprocess_kernel|scatter_kernel|tile_scan' -s 60 -c 5 \
    -o $OUT/prof_split python tools/single_frame.py 256 3 > $OUT/ncu_full_split.log 2>&1
ls -la $OUT
