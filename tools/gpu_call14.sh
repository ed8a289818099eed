mkdir -p gpurun_out
for i in 1 2 3; do
OD_BENCH_DEBUG=1 OD_BENCH_TAG=_r$i timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --no-legs > gpurun_out/t14_bench_$i.json 2> gpurun_out/t14_bench_$i.err
python - <<PY
import json
d = json.loads(open('gpurun_out/t14_bench_$i.json').read().strip().splitlines()[-1])
s = json.load(open('gpurun_out/steps_n1_rank0_r$i.json'))
print('run $i value', d['value'], 'ms/step', d['ms_per_step'], 'total', s['ms_total'])
print('  step_ms', s['step_ms'])
print('  gap_ms', s['gap_ms'])
PY
done
