N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/test_peer_gather.py 2>&1 | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/multi_${N}.err | tail -1 > gpurun_out/multi_${N}.json; cut -c1-300 gpurun_out/multi_${N}.json; python -c "
import json; d=json.loads(open('gpurun_out/multi_${N}.json').read()); print('value',round(d['value']),'e2e',d['e2e'])"
tail -3 gpurun_out/multi_${N}.err
