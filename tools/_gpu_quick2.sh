mkdir -p gpurun_out/bm
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_kernel -s 4 -c 1 -o gpurun_out/bm/fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bm/log.txt 2>&1
ls -la gpurun_out/bm
