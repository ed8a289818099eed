#!/bin/bash
# run on the GPU box: kernel-alone times of the step kernels for every tuning build in build_variants/
for so in build_variants/*.so; do
  ODCUDA_LIB=$PWD/$so python bench.py --steps 12 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-34s fused %.3f ms  current-only %.3f ms  exact %.3f  fast %.3f' % ('$so', d['roofline']['kernel_ms'], d['current_only']['kernel_ms'], d['exact_replay_mode']['kernel_ms'], d['fast_mode']['kernel_ms']))"
done
