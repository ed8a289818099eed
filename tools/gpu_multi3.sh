N=${1:-2}
mkdir -p gpurun_out
run_bench () {
  tag=$1; shift
  env OD_BENCH_DEBUG=1 OD_BENCH_TAG=_$tag "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu --no-parity --no-legs > gpurun_out/multi_n${N}_bench_$tag.json 2> gpurun_out/multi_n${N}_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/multi_n${N}_bench_$tag.json').read().strip().splitlines()[-1])
    print('$tag', 'value', d['value'], 'ms/step', d['ms_per_step'])
    for r in range($N):
        s = json.load(open('gpurun_out/steps_n${N}_rank%d_$tag.json' % r))
        print('  rank', r, 'total', s['ms_total'])
        print('   step_ms', s['step_ms'])
except Exception as ex:
    print('$tag: no bench line', ex)
PY
}
run_bench ctas4 OD_BCAST_CTAS=4
run_bench ctas2 OD_BCAST_CTAS=2
run_bench ctas8 OD_BCAST_CTAS=8
run_bench ctas0 OD_BCAST_CTAS=0
