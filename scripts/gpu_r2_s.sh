#!/bin/bash
# gpurun batch S (round 2, 2 GPUs): the committed state at N = 1 and N = 2 on the same box (weak scaling), the 2-rank
# tests over NCCL, the reference arm under torchrun (rank 0 runs it, rank 1 exits).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_distributed.py -q -m gpu --timeout=500 > gpurun_out/r2s_test_gpu_distributed.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err
timeout 600 $TR --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err
MMB_GRAD_ALLREDUCE=bf16 timeout 600 $TR --master-port 29632 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2s_bench_n2_bf16ar.json 2> gpurun_out/r2s_bench_n2_bf16ar.err
timeout 600 $TR --master-port 29633 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2s_bench_reference_n2.json 2> gpurun_out/r2s_bench_reference_n2.err
tail -n 2 gpurun_out/r2s_test_gpu_distributed.log
for f in r2s_bench_n1 r2s_bench_n2 r2s_bench_n2_bf16ar r2s_bench_reference_n2; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; tail -n 1 gpurun_out/$f.err | cut -c1-200; done
