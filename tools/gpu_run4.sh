#!/bin/bash
mkdir -p gpurun_out
echo "== fused tests" ; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 240 -p no:cacheprovider -k "fused or bf16 or teacher or greedy" 2>&1 | tail -30 > gpurun_out/pytest_fused.log; tail -8 gpurun_out/pytest_fused.log
echo "== phase profile"; timeout 300 python tools/profile_step.py 100 2>&1 | tail -14 | tee gpurun_out/step_phases.txt
echo "== bench fused" ; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1; tail -1 gpurun_out/bench_fused.log | cut -c1-200; grep -o '"roofline".*' gpurun_out/bench_fused.log | cut -c1-400
echo "== ncu full (step kernel)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step -s 20 -c 1 -f -o gpurun_out/step_full python tools/profile_step.py 30 > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
