set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/t2_gputests.log 2>&1
tail -5 gpurun_out/t2_gputests.log
grep -h "cfg[245]" gpurun_out/t2_gputests.log | head -20
