#!/bin/bash
mkdir -p gpurun_out
echo "== phase profile"; timeout 300 python tools/profile_step.py 100 2>&1 | tail -20 | tee gpurun_out/step_phases.txt
echo "== ncu full (step kernel)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step -s 20 -c 1 -f -o gpurun_out/step_full python tools/profile_step.py 30 > gpurun_out/ncu_step.log 2>&1; tail -3 gpurun_out/ncu_step.log
