#!/bin/bash
# gpurun batch B (round 2): ncu evidence for the kernels bench.py times + config-4 capability run + FLAVA/CoCa probes.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2b_build.log 2>&1
# 1. full-set captures of one launch of every hot kernel at the benchmarked shapes (second pass of the script)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn_|add_ln_fwd|ln_bwd' \
   --launch-skip 9 --launch-count 9 -f -o gpurun_out/r2_kernels python scripts/ncu_kernels.py > gpurun_out/r2b_ncu_full.log 2>&1
# 2. launch list of one whole benchmark step (durations only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2300 --launch-count 700 --csv \
   --log-file gpurun_out/r2_launches_bs1024.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline \
   > gpurun_out/r2b_bench_under_ncu.log 2>&1
# 3. ViT-L/14, 4096 pairs per GPU, two-pass recompute (BASELINE config 4 on ONE GPU: same per-GPU work as the 8-GPU run)
timeout 900 python bench.py --config l14 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_l14.json 2> gpurun_out/r2b_bench_l14.err
# 4. probes
timeout 300 python scripts/flava_probe.py > gpurun_out/r2b_flava_probe.out 2>&1
timeout 300 python scripts/coca_probe.py > gpurun_out/r2b_coca_probe.out 2>&1
tail -n 3 gpurun_out/r2b_ncu_full.log; wc -l gpurun_out/r2_launches_bs1024.csv; head -c 1800 gpurun_out/r2b_bench_l14.json; tail -n 3 gpurun_out/r2b_bench_l14.err
tail -n 4 gpurun_out/r2b_flava_probe.out; tail -n 3 gpurun_out/r2b_coca_probe.out
