#!/bin/bash
mkdir -p gpurun_out
echo "== phase profile"; timeout -s KILL 120 python tools/profile_step.py 100 2>&1 | tail -13 | tee gpurun_out/step_phases.txt
echo "== bench N=2"; timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -3 gpurun_out/bench_n2.log | cut -c1-600
echo "== bench N=2 reference arm"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; tail -2 gpurun_out/bench_ref_n2.log | cut -c1-300
