#!/bin/bash
# gpurun batch G (round 2, 1 GPU): the fused single-pass attention backward — parity gate, kernel-alone A/B against the
# two-pass column-split kernel, then the whole step with the faster one.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2g_build.log 2>&1
for ns in 1 2; do
  MMB_ATTN_BWD=fused MMB_ATTN_FUSED_STATS=$ns timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2g_gate_fused_ns$ns.log 2>&1
done
( echo "== colsplit (two passes)"; timeout 300 python scripts/attn_bench.py
  echo "== fused, 1 statistics warp (168 regs)"; MMB_ATTN_BWD=fused MMB_ATTN_FUSED_STATS=1 timeout 300 python scripts/attn_bench.py
  echo "== fused, 2 statistics warps (128 regs, spills)"; MMB_ATTN_BWD=fused MMB_ATTN_FUSED_STATS=2 timeout 300 python scripts/attn_bench.py ) 2>&1 | grep -v watchdog > gpurun_out/r2g_attn_bench.log
# whole step + full-size gradient parity with the fused kernel (only meaningful if the gate passed)
if grep -q " passed" gpurun_out/r2g_gate_fused_ns1.log && ! grep -q "failed" gpurun_out/r2g_gate_fused_ns1.log; then
  MMB_ATTN_BWD=fused timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=600 > gpurun_out/r2g_test_gpu_parity_fused.log 2>&1
  MMB_ATTN_BWD=fused timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_flava.py -q -m gpu --timeout=500 > gpurun_out/r2g_test_misc_fused.log 2>&1
  MMB_ATTN_BWD=fused timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2g_bench_fused.json 2> gpurun_out/r2g_bench_fused.err
fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2g_bench_colsplit.json 2> gpurun_out/r2g_bench_colsplit.err
tail -n 3 gpurun_out/r2g_gate_fused_ns1.log gpurun_out/r2g_gate_fused_ns2.log; cat gpurun_out/r2g_attn_bench.log
for f in gpurun_out/r2g_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 6; done
for f in r2g_bench_fused r2g_bench_colsplit; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; done
