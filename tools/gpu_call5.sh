set -x
mkdir -p gpurun_out
python -c 'from opendrift_b200 import build; build.build()' || exit 1
python -m pytest tests -m gpu -q -s > gpurun_out/t5_gputests.log 2>&1
tail -4 gpurun_out/t5_gputests.log
grep -h "FAILED\|Error" gpurun_out/t5_gputests.log | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/t5_bench.json 2> gpurun_out/t5_bench.err
tail -c 400 gpurun_out/t5_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t5_ref.json 2> gpurun_out/t5_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-parity --no-legs > gpurun_out/t5_ncu_bench.log 2>&1
# full captures with the slim build of the same sources (same kernels, a 5x smaller module: the reports embed the module),
# summarised here; only the summaries travel back
export ODCUDA_LIB=$PWD/build_variants/slim.so
mkdir -p /tmp/ncu
timeout 600 ncu --set full --clock-control none -k regex:^step_kernel -s 2 -c 1 -f -o /tmp/ncu/r2_step_kernel_fused10m python tools/profile_step.py > gpurun_out/t5_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:^step_kernel -s 2 -c 1 -f -o /tmp/ncu/r2_step_kernel_fast10m python tools/profile_step.py 10000000 1 > gpurun_out/t5_ncu_fast.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:mix_ -s 1 -c 1 -f -o /tmp/ncu/r2_mix_kernel_5m python tools/profile_kernels.py mix > gpurun_out/t5_ncu_mix.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:leeway_ -s 1 -c 1 -f -o /tmp/ncu/r2_leeway_kernel_20m python tools/profile_kernels.py leeway > gpurun_out/t5_ncu_leeway.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:analytic_ -s 1 -c 1 -f -o /tmp/ncu/r2_analytic_step_kernel_10m python tools/profile_kernels.py analytic > gpurun_out/t5_ncu_analytic.log 2>&1
unset ODCUDA_LIB
ls -la /tmp/ncu
for f in /tmp/ncu/*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  python profiles/ncu_summary.py $f > gpurun_out/$b.md 2> gpurun_out/$b.err
done
# keep the smallest report as evidence if it fits
ls -S /tmp/ncu/*.ncu-rep | tail -1 | xargs -I{} sh -c 'sz=$(stat -c %s {}); [ $sz -lt 20000000 ] && cp {} gpurun_out/'
du -sh gpurun_out
