# N GPUs of one box: the driver's own launch of the bench (default steps), as at round end
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_n${N}.json 2> gpurun_out/r2b_bench_n${N}.err
echo rc $?
tail -c 400 gpurun_out/r2b_bench_n${N}.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2b_bench_n${N}.json').read().strip().splitlines()[-1])
    print('N=${N} value', d['value'], 'ms/step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], 'comm', json.dumps(d.get('comm'))[:400])
    print('api', json.dumps(d.get('api'))[:400])
except Exception as ex:
    print('no bench line', ex)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2b_ref_n${N}.json 2> gpurun_out/r2b_ref_n${N}.err
echo ref rc $?; tail -c 300 gpurun_out/r2b_ref_n${N}.json
