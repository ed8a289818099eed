set -x
mkdir -p gpurun_out
rm -f gpurun_out/t10_kbench.jsonl
timeout 300 python tools/kbench.py >> gpurun_out/t10_kbench.jsonl 2>> gpurun_out/t10_kbench.err
for so in build_variants/*.so; do ODCUDA_LIB=$PWD/$so timeout 300 python tools/kbench.py >> gpurun_out/t10_kbench.jsonl 2>> gpurun_out/t10_kbench.err; done
python - <<'PY'
import json
for l in open('gpurun_out/t10_kbench.jsonl'):
    d = json.loads(l)
    print(d['lib'].split('/')[-1], {k: v for k, v in d.items() if k.endswith('_ms') and 'min' not in k})
    print('   same bits:', [d[k + '_sha'] == d[k + '_gen_sha'] for k in ('fused', 'fused_t1', 'fused_t5', 'cur')])
PY
