#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1"; shift; timeout -s KILL "$@" 2>&1 | tail -4; }
run "ubench" 60 ./tools/ubench_bin
echo "== phase profile"; timeout -s KILL 120 python tools/profile_step.py 100 2>&1 | tail -13 | tee gpurun_out/step_phases.txt
echo "== ncu full step"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:decode_step -s 20 -c 1 -f -o gpurun_out/step_full python tools/profile_step.py 30 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
echo "== ncu full dac"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 150 -c 2 -f -o gpurun_out/dac_full python tools/bench_dac.py 8 248 > gpurun_out/ncu_dac.log 2>&1; tail -1 gpurun_out/ncu_dac.log | cut -c1-200
echo "== ncu launch list (bench)"; timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --decode-steps 64 > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
