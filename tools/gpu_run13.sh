#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout -s KILL "$@" 2>&1 | tail -12; }
run "all gpu tests" 500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 150
echo "== phase profile"; timeout -s KILL 120 python tools/profile_step.py 100 2>&1 | tail -13 | tee gpurun_out/step_phases.txt
