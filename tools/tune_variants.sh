#!/bin/bash
# run on the GPU box: kernel-alone time of step_kernel<RK4> for every tuning build in build_variants/
for so in build_variants/*.so; do
  ODCUDA_LIB=$PWD/$so python bench.py --steps 6 --warmup 3 --no-cpu --sort-every 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-34s kernel %.3f ms  value %.3e  ms/step %.3f' % ('$so', d['roofline']['kernel_ms'], d['value'], d['ms_per_step']))"
done
