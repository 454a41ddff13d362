#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout -s KILL 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider --timeout 150 2>&1 | tail -8
echo "== phase profile"; timeout -s KILL 120 python tools/profile_step.py 100 2>&1 | tail -13 | tee gpurun_out/step_phases.txt
