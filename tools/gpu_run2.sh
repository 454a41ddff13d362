#!/bin/bash
mkdir -p gpurun_out
echo "== fused tests" ; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 240 -p no:cacheprovider -k "fused or bf16" 2>&1 | tail -60 > gpurun_out/pytest_fused.log; tail -30 gpurun_out/pytest_fused.log
echo "== bench fused" ; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1; tail -2 gpurun_out/bench_fused.log
echo "== pytest all" ; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest.log; tail -15 gpurun_out/pytest.log
