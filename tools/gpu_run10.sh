#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout -s KILL "$@" 2>&1 | tail -6; }
run "dac tests" 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "dac" --timeout 100
echo "== dac bench"; timeout -s KILL 200 python tools/bench_dac.py 32 248 2>&1 | tail -2
echo "== bench full"; timeout -s KILL 500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-1500
