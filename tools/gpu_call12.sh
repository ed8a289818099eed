mkdir -p gpurun_out
timeout 600 python tools/leg_probe.py > gpurun_out/t12_leg_probe.json 2> gpurun_out/t12_leg_probe.err
cat gpurun_out/t12_leg_probe.json; tail -3 gpurun_out/t12_leg_probe.err
