#!/bin/bash
mkdir -p gpurun_out
echo "== dac bench"; timeout -s KILL 200 python tools/bench_dac.py 32 248 2>&1 | tail -3
echo "== ncu dac launches"; timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc -s 120 -c 60 --csv --log-file gpurun_out/dac_tc_launches.csv python tools/bench_dac.py 8 248 > gpurun_out/ncu_dac.log 2>&1; tail -2 gpurun_out/ncu_dac.log | cut -c1-300
