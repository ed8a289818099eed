set -x
mkdir -p gpurun_out
N=${1:-2}
python -c 'from opendrift_b200 import build; build.build()' || exit 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/nccl_tile_run.py > gpurun_out/multi_n${N}_tiles.jsonl 2> gpurun_out/multi_n${N}_tiles.err
cat gpurun_out/multi_n${N}_tiles.jsonl
tail -c 1500 gpurun_out/multi_n${N}_tiles.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu > gpurun_out/multi_n${N}_bench.json 2> gpurun_out/multi_n${N}_bench.err
tail -c 800 gpurun_out/multi_n${N}_bench.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/multi_n${N}_bench.json').read().strip().splitlines()[-1])
    for k in ('value', 'ms_per_step', 'n_gpus', 'comm', 'e2e', 'parity'):
        print(k, json.dumps(d.get(k))[:900])
    print(d['config']['parallelism'])
except Exception as ex:
    print('no bench line', ex)
PY
