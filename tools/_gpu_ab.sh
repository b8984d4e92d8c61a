for v in c4; do
  if [ -z "$v" ]; then unset GOF_B200_LIB; name=base; else export GOF_B200_LIB=/root/repo/f3d_gaus_b200/variants/libgof_b200_$v.so; name=$v; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 20 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$name','value',round(d['value']),'e2e',round(d['e2e']['value']),'per_view',round(d['per_view_api']['value']),'stages',{k:round(v*1e3) for k,v in r['stage_ms_per_step'].items()})
"
done
