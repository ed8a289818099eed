set -x
mkdir -p gpurun_out
python -c 'from opendrift_b200 import build; build.build()' || exit 1
python -m pytest tests -m gpu -q -s > gpurun_out/t4_gputests.log 2>&1
tail -4 gpurun_out/t4_gputests.log
grep -h "FAILED\|Error" gpurun_out/t4_gputests.log | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/t4_bench.json 2> gpurun_out/t4_bench.err
tail -c 400 gpurun_out/t4_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t4_ref.json 2> gpurun_out/t4_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-parity --no-legs > gpurun_out/t4_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:^step_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_step_final python tools/profile_step.py > gpurun_out/t4_ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:^step_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_step_fast python tools/profile_step.py 10000000 1 > gpurun_out/t4_ncu_fast.log 2>&1
for k in mix leeway analytic; do
  timeout 600 ncu --set full --clock-control none -k regex:${k}_ -s 1 -c 1 -f -o gpurun_out/prof_r2_${k} python tools/profile_kernels.py $k > gpurun_out/t4_ncu_${k}.log 2>&1
done
ls -la gpurun_out/prof_r2_*
du -sh gpurun_out
