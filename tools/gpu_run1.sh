#!/bin/bash
# First GPU pass: parity tests, smoke, a short bench, and the per-launch ncu timing list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest" ; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/pytest.log; tail -40 gpurun_out/pytest.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
echo "== bench nopdl" ; PTTS_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.log 2>&1; tail -2 gpurun_out/bench_nopdl.log
echo "== ncu launches" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 420 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --decode-steps 24 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
