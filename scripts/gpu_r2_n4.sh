#!/bin/bash
# gpurun batch N4 (round 2, 4 GPUs): the committed state at N = 4 (weak scaling) — one bench line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29651 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2n4_bench_n4.json 2> gpurun_out/r2n4_bench_n4.err
grep '^{' gpurun_out/r2n4_bench_n4.json | head -c 400; echo; tail -n 2 gpurun_out/r2n4_bench_n4.err | cut -c1-300
