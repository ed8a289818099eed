set -x
mkdir -p gpurun_out
rm -f gpurun_out/t8_kbench.jsonl
for so in build_variants/*.so; do ODCUDA_LIB=$PWD/$so timeout 300 python tools/kbench.py >> gpurun_out/t8_kbench.jsonl 2>> gpurun_out/t8_kbench.err; done
python - <<'PY'
import json
for l in open('gpurun_out/t8_kbench.jsonl'):
    d = json.loads(l)
    print('%-40s fused %.4f gen %.4f %s | t1 %.4f gen %.4f %s | cur %.4f gen %.4f %s' % (d['lib'].split('/')[-1], d['fused_ms'], d['fused_gen_ms'], d['fused_sha'] == d['fused_gen_sha'],
          d['fused_t1_ms'], d['fused_t1_gen_ms'], d['fused_t1_sha'] == d['fused_t1_gen_sha'], d['cur_ms'], d['cur_gen_ms'], d['cur_sha'] == d['cur_gen_sha']))
PY
