#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout -s KILL "$@" 2>&1 | tail -25; }
run "dac tests" 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "dac" --timeout 100
