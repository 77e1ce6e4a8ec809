#!/bin/bash
# gpurun batch N8 (round 2, 8 GPUs): the committed state at N = 8 (weak scaling) — one bench line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 700 $TR --master-port 29641 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n8_bench_n8.json 2> gpurun_out/r2n8_bench_n8.err
grep '^{' gpurun_out/r2n8_bench_n8.json | head -c 1200; echo; tail -n 3 gpurun_out/r2n8_bench_n8.err | cut -c1-300
