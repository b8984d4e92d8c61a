mkdir -p gpurun_out/ab2
for v in slab96 slab80; do
  export GOF_B200_LIB=/root/repo/f3d_gaus_b200/variants/libgof_b200_$v.so
  timeout 600 ncu --set full --clock-control none -k regex:render_fwd_kernel -s 4 -c 1 -o gpurun_out/ab2/$v python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab2/$v.log 2>&1
done
ls -la gpurun_out/ab2
