#!/bin/bash
# gpurun batch K (round 2, 1 GPU): anti-phase start of the item forward kernel (A/B), packed-f32x2 math in the fused
# backward workers.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2k_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2k_gate_default.log 2>&1
MMB_ATTN_FWD=item timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2k_gate_item.log 2>&1
( echo "== fwd tile (default), bwd fused f32x2"; timeout 300 python scripts/attn_bench.py
  echo "== fwd item, anti-phase start"; MMB_ATTN_FWD=item timeout 300 python scripts/attn_bench.py
  echo "== fwd item, lockstep start"; MMB_ATTN_FWD=item MMB_ATTN_ITEM_STAGGER=0 timeout 300 python scripts/attn_bench.py ) 2>&1 | grep -v watchdog > gpurun_out/r2k_attn_bench.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flava.py tests/test_gpu_distributed.py -q -m gpu --timeout=600 > gpurun_out/r2k_test_gpu.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_default.err
MMB_ATTN_FWD=item timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2k_bench_item.json 2> gpurun_out/r2k_bench_item.err
tail -n 4 gpurun_out/r2k_gate_default.log gpurun_out/r2k_gate_item.log; cat gpurun_out/r2k_attn_bench.log
for f in gpurun_out/r2k_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 6; done
for f in r2k_bench_default r2k_bench_item; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; done
