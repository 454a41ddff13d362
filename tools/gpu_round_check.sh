#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout -s KILL 240 python -u -m pytest tests -v -m gpu -x -p no:cacheprovider --timeout 100 2>&1 | tee gpurun_out/tests.log | tail -3 | cut -c1-150
echo "== smoke"; timeout -s KILL 120 python -u -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== phase profile"; timeout -s KILL 100 python -u tools/profile_step.py 100 2>&1 | tail -14 | tee gpurun_out/step_phases.txt
echo "== bench"; timeout -s KILL 300 python -u bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
