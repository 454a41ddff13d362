#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout -s KILL 90 "$@" 2>&1 | tail -12; }
run "legacy bf16 (PTTS_FUSED=0)" env PTTS_FUSED=0 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "teacher_forced_bf16 and abs"
run "fused bf16" env PTTS_FUSED=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "teacher_forced_bf16 and abs"
