#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout -s KILL "$@" 2>&1 | tail -6; }
run "fused tests" 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "fused or bf16" --timeout 150
echo "== phase profile"; timeout -s KILL 120 python tools/profile_step.py 100 2>&1 | tail -13 | tee gpurun_out/step_phases.txt
