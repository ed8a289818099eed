set -x
mkdir -p gpurun_out
# the library must match the sources of this snapshot (rebuilds only when a source is newer than the .so)
python -c 'from opendrift_b200 import build; build.build()' || exit 1
python -m pytest tests -m gpu -x -q -s > gpurun_out/t2_gputests.log 2>&1
tail -5 gpurun_out/t2_gputests.log
grep -h "cfg[245]" gpurun_out/t2_gputests.log | head -20
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/t2_bench.json 2> gpurun_out/t2_bench.err
tail -c 1500 gpurun_out/t2_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/t2_bench.json').read().strip().splitlines()[-1])
    for k in ('value', 'ms_per_step', 'e2e', 'parity', 'api', 'cfg4_mixing_wind_stokes', 'cfg5_leeway', 'roofline'):
        print(k, json.dumps(d.get(k))[:1500])
except Exception as ex:
    print('no bench line', ex)
PY
