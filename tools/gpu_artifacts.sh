#!/bin/bash
# Round artifacts: bench line, ncu launch list (own kernels only), ncu --set full of the decode-step kernel and of the DAC conv kernels.
# Large .ncu-rep files are reduced to CSV on the box (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
K='regex:decode_step|conv_tc_kernel|final_conv|linear_|attention|embed_kernel|sample|from_codes|conv_kernel|relayout|gather|row_stats'
export PTTS_STEP_COOP=0   # Nsight Compute cannot launch a cooperative cluster grid
export PTTS_STEPS_PER_LAUNCH=1   # one token per launch: the replays of a 64-token launch would take minutes, and `traffic` is per token
echo "== ncu launch list (bench command, 64 decode steps)"
timeout -s KILL 700 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1300 --csv --log-file gpurun_out/launches.csv python -u bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --decode-steps 64 > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches.csv
echo "== ncu full step"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:decode_step -s 40 -c 1 -o gpurun_out/step_full -f python -u bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-dac --decode-steps 64 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log | cut -c1-200
ncu -i gpurun_out/step_full.ncu-rep --page raw --csv > gpurun_out/step_full_raw.csv 2>/dev/null
ncu -i gpurun_out/step_full.ncu-rep --page details --csv > gpurun_out/step_full_details.csv 2>/dev/null
rm -f gpurun_out/step_full.ncu-rep
echo "== ncu full dac"
PTTS_DAC_BENCH_TC_ONLY=1 timeout -s KILL 300 ncu --set full --clock-control none -k regex:"conv_tc_kernel|conv_kernel|final_conv|from_codes" -s 64 -c 36 -o gpurun_out/dac_full -f python -u tools/bench_dac.py > gpurun_out/ncu_dac.log 2>&1; tail -1 gpurun_out/ncu_dac.log | cut -c1-200
ncu -i gpurun_out/dac_full.ncu-rep --page raw --csv > gpurun_out/dac_full_raw.csv 2>/dev/null; rm -f gpurun_out/dac_full.ncu-rep
echo "== SASS mnemonics of the hot kernels"
for k in decode_step_cluster_kernel decode_step_kernel conv_tc_kernel linear_tc_kernel; do
  echo "-- $k"; cuobjdump -sass parler_tts_b200/csrc/libptts_b200.so 2>/dev/null | awk -v k="$k" '/Function :/{f=($0 ~ k)} f' | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | grep -E "HMMA|UTC|LDTM|UBLK|UTMA|LDSM|SYNCS|UCGABAR|BAR|MUFU|RED|ATOM" | sort | uniq -c | sort -rn | head -14
done > gpurun_out/sass_mnemonics.txt 2>&1; wc -l gpurun_out/sass_mnemonics.txt
du -sh gpurun_out
