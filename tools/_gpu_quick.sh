# quick iteration: forward/batch parity + our bench arm only
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -q -k "forward or batch or workspace or render_views or big_tile" 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --steps 20 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'per_view',round(d['per_view_api']['value']),round(d['per_view_api']['e2e']))
print('stages',{k:round(v*1e3) for k,v in r['stage_ms_per_step'].items()},'frac',round(r['frac'],4))
"
