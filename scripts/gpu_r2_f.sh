#!/bin/bash
# gpurun batch F (round 2, 1 GPU): checkpoint of the committed state — every GPU test file, smoke, bench (headline line),
# launch list of one step, refreshed ncu captures, probes.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2300 --launch-count 700 --csv \
   --log-file gpurun_out/r2_launches_bs1024.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline \
   > gpurun_out/r2f_bench_under_ncu.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn_|add_ln_fwd|ln_bwd' \
   --launch-skip 9 --launch-count 9 -f -o gpurun_out/r2_kernels_final python scripts/ncu_kernels.py > gpurun_out/r2f_ncu_full.log 2>&1
timeout 300 python scripts/flava_probe.py > gpurun_out/r2f_flava_probe.out 2>&1
timeout 300 python scripts/coca_probe.py > gpurun_out/r2f_coca_probe.out 2>&1
tail -n 2 gpurun_out/r2f_smoke.log; grep -E "^FAILED|passed|failed" gpurun_out/r2f_pytest_gpu.log | tail -n 8
grep '^{' gpurun_out/r2f_bench.json | head -c 1200; echo; grep '^{' gpurun_out/r2f_bench_reference.json | head -c 600; echo
wc -l gpurun_out/r2_launches_bs1024.csv; tail -n 2 gpurun_out/r2f_ncu_full.log; tail -n 3 gpurun_out/r2f_flava_probe.out; tail -n 2 gpurun_out/r2f_coca_probe.out
