#!/bin/bash
# A/B of kernel variants on the headline workload: tools/ab.sh <out-dir> <label>=<lib-or-empty>[,ENV=VAL...] ...
OUT=$1; shift; mkdir -p $OUT
for spec in "$@"; do
  label=${spec%%=*}; rest=${spec#*=}
  lib=${rest%%,*}; envs=""
  if [[ "$rest" == *,* ]]; then envs=$(echo "${rest#*,}" | tr ',' ' '); fi
  if [ -n "$lib" ]; then envs="$envs GOF_B200_LIB=$lib"; fi
  env $envs timeout 300 python bench.py --no-others --no-cpu-baseline --steps 200 > $OUT/ab_$label.json 2> $OUT/ab_$label.err
  python - "$OUT/ab_$label.json" "$label" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"serial",round(d["e2e_serial"]["value"]),"per_view",round(d["per_view_api"]["value"]),round(d["per_view_api"]["e2e"]),
          "stages_us",{k:round(v*1e3,1) for k,v in r["stage_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
