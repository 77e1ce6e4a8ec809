#!/bin/bash
# gpurun batch C (round 2): attention A/B (column-split +/- prefetch, ping-pong with deep ring), remaining test fixes,
# bench with the measured-best variants, ncu evidence, config-4 capability run.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2c_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2c_gate_attn.log 2>&1
MMB_ATTN_BWD=pp timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2c_gate_attn_pp.log 2>&1
( echo "== default (colsplit + prefetch; fwd: pp for S<=128, tile otherwise)"; timeout 300 python scripts/attn_bench.py
  echo "== colsplit, no prefetch"; MMB_ATTN_PREFETCH=0 timeout 300 python scripts/attn_bench.py
  echo "== bwd pp (ring 6/5)"; MMB_ATTN_BWD=pp timeout 300 python scripts/attn_bench.py ) 2>&1 | grep -v watchdog > gpurun_out/r2c_attn_bench.log
for f in tests/test_gpu_parity.py tests/test_gpu_layers.py; do
  timeout 900 python -m pytest $f -q -m gpu --timeout=600 > gpurun_out/r2c_$(basename $f .py).log 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
# ncu: full-set captures of one launch of every hot kernel at the benchmarked shapes (second pass of the script)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn_|add_ln_fwd|ln_bwd' \
   --launch-skip 9 --launch-count 9 -f -o gpurun_out/r2_kernels python scripts/ncu_kernels.py > gpurun_out/r2c_ncu_full.log 2>&1
# ViT-L/14, 4096 pairs per GPU, two-pass recompute (BASELINE config 4 per-GPU work on ONE GPU)
timeout 900 python bench.py --config l14 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_l14.json 2> gpurun_out/r2c_bench_l14.err
cat gpurun_out/r2c_attn_bench.log; tail -n 2 gpurun_out/r2c_gate_attn.log gpurun_out/r2c_gate_attn_pp.log
for f in gpurun_out/r2c_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 6; done
head -c 600 gpurun_out/r2c_bench.json; echo; tail -n 3 gpurun_out/r2c_ncu_full.log; head -c 1500 gpurun_out/r2c_bench_l14.json; tail -n 3 gpurun_out/r2c_bench_l14.err
