mkdir -p gpurun_out
for c in 12 24 48 96; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-parity --no-legs --e2e-chunks $c > gpurun_out/t16_bench_c$c.json 2> gpurun_out/t16_bench_c$c.err
python - <<PY
import json
d = json.loads(open('gpurun_out/t16_bench_c$c.json').read().strip().splitlines()[-1])
print('chunks $c value', d['value'], 'e2e', d['e2e']['value'], 'pcie', d['e2e']['pcie_probe'], 'pageable', d['e2e'].get('pageable_numpy', {}).get('value'))
PY
done
