set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t1_gputests.log 2>&1
rm -f gpurun_out/t1_kbench.jsonl
for so in build_variants/*.so; do ODCUDA_LIB=$PWD/$so timeout 300 python tools/kbench.py >> gpurun_out/t1_kbench.jsonl 2>> gpurun_out/t1_kbench.err; done
timeout 300 python tools/kbench.py >> gpurun_out/t1_kbench.jsonl 2>> gpurun_out/t1_kbench.err
cat gpurun_out/t1_kbench.jsonl
ODCUDA_LIB=$PWD/build_variants/s2_ord3_mb8.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_s2mb8 python tools/profile_step.py > gpurun_out/t1_ncu8.log 2>&1
ODCUDA_LIB=$PWD/build_variants/s2_ord3_mb6.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_s2mb6 python tools/profile_step.py > gpurun_out/t1_ncu6.log 2>&1
tail -3 gpurun_out/t1_gputests.log
