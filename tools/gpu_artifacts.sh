#!/bin/bash
# Round artifacts: bench line, ncu launch list (own kernels only), ncu --set full of the step kernel and of the DAC conv kernels.
# Large .ncu-rep files are reduced to CSV on the box (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
echo "== bench"; timeout -s KILL 400 python -u bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
echo "== ncu launch list"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_step_kernel|conv_tc_kernel|linear_|attention|embed|sample|from_codes|dac_|relayout|gather" -c 700 --csv --log-file gpurun_out/launches.csv python -u bench.py --steps 1 --warmup 1 --no-cpu-baseline --decode-steps 64 > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches.csv
echo "== ncu full step"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 40 -c 1 -o gpurun_out/step_full -f python -u bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-dac --decode-steps 64 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log | cut -c1-200
ncu -i gpurun_out/step_full.ncu-rep --page raw --csv > gpurun_out/step_full_raw.csv 2>/dev/null
echo "== ncu full dac"
timeout -s KILL 200 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 30 -c 30 -o gpurun_out/dac_full -f python -u tools/bench_dac.py > gpurun_out/ncu_dac.log 2>&1; tail -1 gpurun_out/ncu_dac.log | cut -c1-200
ncu -i gpurun_out/dac_full.ncu-rep --page raw --csv > gpurun_out/dac_full_raw.csv 2>/dev/null; rm -f gpurun_out/dac_full.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | head -20
