#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout -s KILL 240 python -u -m pytest tests -v -m gpu -x -p no:cacheprovider --timeout 100 2>&1 | tee gpurun_out/tests.log | tail -12 | cut -c1-150
echo "== phase profile"; timeout -s KILL 100 python -u tools/profile_step.py 100 2>&1 | tail -13 | tee gpurun_out/step_phases.txt
