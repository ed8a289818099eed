set -x
mkdir -p gpurun_out
N=${1:-2}
python -c 'from opendrift_b200 import build; build.build()' || exit 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/nccl_tile_run.py > gpurun_out/multi_n${N}_tiles.jsonl 2> gpurun_out/multi_n${N}_tiles.err
cat gpurun_out/multi_n${N}_tiles.jsonl | cut -c1-700
tail -c 600 gpurun_out/multi_n${N}_tiles.err
run_bench () {   # tag, env...
  tag=$1; shift
  env OD_BENCH_DEBUG=1 OD_BENCH_TAG=_$tag "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu --no-parity --no-legs > gpurun_out/multi_n${N}_bench_$tag.json 2> gpurun_out/multi_n${N}_bench_$tag.err
  tail -c 300 gpurun_out/multi_n${N}_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/multi_n${N}_bench_$tag.json').read().strip().splitlines()[-1])
    print('$tag', 'value', d['value'], 'ms/step', d['ms_per_step'], 'comm', json.dumps(d.get('comm'))[:300], 'e2e', d['e2e']['value'])
    for r in range($N):
        s = json.load(open('gpurun_out/steps_n${N}_rank%d_$tag.json' % r))
        big = [(k, v) for k, v in enumerate(s['step_ms']) if v > 1.5]
        gaps = [(k, v) for k, v in enumerate(s['gap_ms']) if v > 0.15]
        print('  rank', r, 'total', s['ms_total'], 'slow steps', big[:12], 'gaps', gaps[:12])
except Exception as ex:
    print('$tag: no bench line', ex)
PY
}
run_bench default
run_bench noprefetch OD_BENCH_NOPREFETCH=1
run_bench nofill OD_BENCH_NOFILL=1
