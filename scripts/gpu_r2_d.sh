#!/bin/bash
# gpurun batch D (round 2, 2 GPUs): the 2-rank tests over NCCL + real peer mappings, and the N=2 benchmark line with the
# non-overlapped (default) and the round-1 overlapped gradient all-reduce schedule.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2d_build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2d_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_distributed.py -q -m gpu -s --timeout=500 > gpurun_out/r2d_test_gpu_distributed.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
   bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
MMB_OVERLAP_ALLREDUCE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
   bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_n2_overlap.json 2> gpurun_out/r2d_bench_n2_overlap.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
tail -n 4 gpurun_out/r2d_test_gpu_distributed.log
for f in r2d_bench_n1 r2d_bench_n2 r2d_bench_n2_overlap; do echo "== $f"; head -c 420 gpurun_out/$f.json; echo; tail -n 2 gpurun_out/$f.err; done
