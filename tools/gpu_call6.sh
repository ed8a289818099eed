set -x
mkdir -p gpurun_out
python -c 'from opendrift_b200 import build; build.build()' || exit 1
python -m pytest tests -m gpu -q -s > gpurun_out/t6_gputests.log 2>&1
tail -4 gpurun_out/t6_gputests.log
grep -h "FAILED\|Error" gpurun_out/t6_gputests.log | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/t6_bench.json 2> gpurun_out/t6_bench.err
tail -c 400 gpurun_out/t6_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/t6_bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'fast', d['fast_mode']['kernel_ms'], 'cur', d['current_only']['kernel_ms'])
    print('api', d['api']['output_at_end']['ms_per_step_steady'], d['api']['output_every_step']['ms_per_step_steady'])
    c4, c5 = d['cfg4_mixing_wind_stokes'], d['cfg5_leeway']
    print('cfg4', c4['ms_per_step'], c4['mix_kernel']['kernel_ms'], c4['step_kernel_all_extras']['kernel_ms'], 'cfg5', c5['ms_per_step'], c5['leeway_kernel']['kernel_ms'])
    print('parity', d['parity']['ok'], c4['parity']['ok'], c5['parity']['ok'])
except Exception as ex:
    print('no bench line', ex)
PY
du -sh gpurun_out
