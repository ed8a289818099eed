set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
# spec-vs-general equality on the GPU (full library)
timeout 900 python -m pytest tests/test_zz_gpu_spec.py -q -x > gpurun_out/t7_spec_tests.log 2>&1
tail -5 gpurun_out/t7_spec_tests.log
# kernel-alone timings of every tuning build
rm -f gpurun_out/t7_kbench.jsonl
for so in build_variants/*.so; do ODCUDA_LIB=$PWD/$so timeout 300 python tools/kbench.py >> gpurun_out/t7_kbench.jsonl 2>> gpurun_out/t7_kbench.err; done
cat gpurun_out/t7_kbench.jsonl
# instruction counts / pipes of the specialised and the general kernel
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
ODCUDA_LIB=$PWD/build_variants/specnc.so timeout 600 ncu --metrics $M --clock-control none -k regex:^step_ -s 2 -c 2 --csv --log-file gpurun_out/t7_ncu_spec.csv python tools/profile_step.py > gpurun_out/t7_ncu_spec.log 2>&1
ODCUDA_LIB=$PWD/build_variants/specnc_mb6.so timeout 600 ncu --metrics $M --clock-control none -k regex:^step_ -s 2 -c 2 --csv --log-file gpurun_out/t7_ncu_spec_mb6.csv python tools/profile_step.py > gpurun_out/t7_ncu_spec6.log 2>&1
grep -h "step_" gpurun_out/t7_ncu_spec.csv | cut -d, -f5,13- | head -40
