#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout -s KILL "$@" 2>&1 | tail -8; }
run "fused bf16 quick" 90 env PTTS_FUSED=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "teacher_forced_bf16"
run "all gpu tests" 400 python -m pytest tests -q -m gpu -x -p no:cacheprovider --timeout 120
echo "== phase profile"; timeout -s KILL 120 python tools/profile_step.py 100 2>&1 | tail -14 | tee gpurun_out/step_phases.txt
echo "== bench fused" ; timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1; tail -1 gpurun_out/bench_fused.log | cut -c1-200; grep -o '"roofline".*' gpurun_out/bench_fused.log | cut -c1-400
