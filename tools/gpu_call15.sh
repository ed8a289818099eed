set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/t15_gputests.log 2>&1
tail -4 gpurun_out/t15_gputests.log
grep -h "FAILED\|Error" gpurun_out/t15_gputests.log | head -20
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/t15_bench.json 2> gpurun_out/t15_bench.err
tail -c 400 gpurun_out/t15_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/t15_bench.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('value', d['value'], d['ms_per_step'], 'kernel', r['kernel_ms'], r.get('kernel_ms_by_alignment'), 'general', r.get('general_kernel_ms'), 'frac', r['frac'], 'fast', d['fast_mode']['kernel_ms'], 'cur', d['current_only']['kernel_ms'])
    print('e2e', d['e2e']['value'])
    print('api', d['api']['output_at_end']['ms_per_step_steady'], d['api']['output_every_step']['ms_per_step_steady'])
    c4, c5 = d['cfg4_mixing_wind_stokes'], d['cfg5_leeway']
    print('cfg4', c4['ms_per_step'], c4['mix_kernel']['kernel_ms'], c4['step_kernel_all_extras']['kernel_ms'], 'cfg5', c5['ms_per_step'], c5['leeway_kernel']['kernel_ms'])
    print('parity', d['parity']['ok'], c4['parity']['ok'], c5['parity']['ok'])
except Exception as ex:
    print('no bench line', ex)
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t15_ref.json 2> gpurun_out/t15_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-parity --no-legs > gpurun_out/t15_ncu_bench.log 2>&1
du -sh gpurun_out
