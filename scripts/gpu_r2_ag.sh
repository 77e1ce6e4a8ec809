#!/bin/bash
# gpurun batch AG (round 2, 1 GPU): one polling lane per warp in mbarrier waits (fewer smem polls beside the MMAs).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2ag_smoke.log 2>&1; tail -n 1 gpurun_out/r2ag_smoke.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd or gemm" --timeout=200 > gpurun_out/r2ag_gate.log 2>&1; tail -n 2 gpurun_out/r2ag_gate.log
timeout 300 python scripts/attn_bench.py 2>&1 | grep -v watchdog > gpurun_out/r2ag_attn_bench.log; cat gpurun_out/r2ag_attn_bench.log
timeout 600 python scripts/gemm_bench.py 2>&1 | head -12 > gpurun_out/r2ag_gemm_bench.log; cat gpurun_out/r2ag_gemm_bench.log | cut -c1-100
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err
grep '^{' gpurun_out/r2ag_bench.json | head -c 300; echo
