timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --steps 20 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'per_view',round(d['per_view_api']['value']),round(d['per_view_api']['e2e']))
print('stages',{k:round(v*1e3) for k,v in r['stage_ms_per_step'].items()},'frac',round(r['frac'],4))
"
timeout 300 python tools/quick_bench.py 256 256 50
