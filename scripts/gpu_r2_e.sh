#!/bin/bash
# gpurun batch E (round 2, 2 GPUs): where do the N=2 milliseconds go — all-reduce alone (fp32 / bf16), the N=2 bench with
# fp32 and bf16-compressed gradient all-reduce, N=1 on the same box.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2e_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29621 scripts/allreduce_probe.py > gpurun_out/r2e_allreduce_probe.out 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
timeout 600 $TR --master-port 29622 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
MMB_GRAD_ALLREDUCE=bf16 timeout 600 $TR --master-port 29623 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_n2_bf16ar.json 2> gpurun_out/r2e_bench_n2_bf16ar.err
timeout 300 python scripts/attn_bench.py 2>&1 | grep -v watchdog > gpurun_out/r2e_attn_bench.log
cat gpurun_out/r2_allreduce_probe_2.log; cat gpurun_out/r2e_attn_bench.log
for f in r2e_bench_n1 r2e_bench_n2 r2e_bench_n2_bf16ar; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; tail -n 1 gpurun_out/$f.err | cut -c1-200; done
