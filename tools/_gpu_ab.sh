for v in "" bwd3; do
  if [ -z "$v" ]; then unset GOF_B200_LIB; name=base; else export GOF_B200_LIB=/root/repo/f3d_gaus_b200/variants/libgof_b200_$v.so; name=$v; fi
  echo $name; timeout 300 python tools/quick_bench.py 256 256 50 | grep ours
done
