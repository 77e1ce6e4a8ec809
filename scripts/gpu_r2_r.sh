#!/bin/bash
# gpurun batch R (round 2, 1 GPU): evidence of the current state — smoke, every GPU test file, bench (both arms),
# launch list of one step, full ncu capture of the hot kernels, probes, the ViT-L/14 configuration.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2r_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/r2r_pytest_gpu.log 2>&1; tail -n 2 gpurun_out/r2r_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2200 --launch-count 700 --csv \
   --log-file gpurun_out/r2r_launches_bs1024.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline \
   > gpurun_out/r2r_bench_under_ncu.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn_|add_ln_fwd|ln_bwd' \
   --launch-skip 8 --launch-count 8 -f -o gpurun_out/r2_kernels_final2 python scripts/ncu_kernels.py > gpurun_out/r2r_ncu_full.log 2>&1
timeout 120 ./build/softmax_probe > gpurun_out/r2r_softmax_probe.log 2>&1
timeout 200 python scripts/probes/hbm_probe.py > gpurun_out/r2r_hbm_probe.log 2>&1
timeout 200 python scripts/attn_item_trace.py > gpurun_out/r2r_item_trace.log 2>&1
timeout 900 python bench.py --config l14 --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2r_bench_l14.json 2> gpurun_out/r2r_bench_l14.err
grep '^{' gpurun_out/r2r_bench.json | head -c 900; echo
grep '^{' gpurun_out/r2r_bench_reference.json | head -c 400; echo
wc -l gpurun_out/r2r_launches_bs1024.csv; tail -n 2 gpurun_out/r2r_ncu_full.log
head -n 12 gpurun_out/r2r_softmax_probe.log; cat gpurun_out/r2r_hbm_probe.log; head -n 9 gpurun_out/r2r_item_trace.log | cut -c1-300
grep '^{' gpurun_out/r2r_bench_l14.json | head -c 400; echo
