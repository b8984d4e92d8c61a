#!/bin/bash
# Multi-GPU bench of both arms: tools/gpu_multi.sh <N> [out-dir] [extra bench flags...]
N=${1:-2}; OUT=${2:-gpurun_out}; shift; shift
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 "$@" 2>$OUT/multi_${N}.err | tail -1 > $OUT/multi_${N}.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 30 --warmup 5 "$@" 2>$OUT/multi_ref_${N}.err | tail -1 > $OUT/multi_ref_${N}.json
python - $OUT/multi_${N}.json $OUT/multi_ref_${N}.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"],3),
              "serial", round(d["e2e_serial"]["value"]) if d.get("e2e_serial") else None, (d.get("e2e_serial") or {}).get("exchange_checked"), d["e2e"].get("readback"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
