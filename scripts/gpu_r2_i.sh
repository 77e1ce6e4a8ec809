#!/bin/bash
# gpurun batch I (round 2, 1 GPU): the item-level forward attention kernel (K/V once per head, two MMA issuers, packed
# f32x2 softmax) — parity gate, kernel-alone A/B against the tile kernel, then the whole step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2i_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2i_gate_default.log 2>&1
MMB_ATTN_FWD=item timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2i_gate_item.log 2>&1
( echo "== fwd tile (default)"; timeout 300 python scripts/attn_bench.py
  echo "== fwd item"; MMB_ATTN_FWD=item timeout 300 python scripts/attn_bench.py ) 2>&1 | grep -v watchdog > gpurun_out/r2i_attn_bench.log
if grep -q " passed" gpurun_out/r2i_gate_item.log && ! grep -q "failed" gpurun_out/r2i_gate_item.log; then
  MMB_ATTN_FWD=item timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flava.py -q -m gpu --timeout=600 > gpurun_out/r2i_test_gpu_item.log 2>&1
  MMB_ATTN_FWD=item timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2i_bench_item.json 2> gpurun_out/r2i_bench_item.err
fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err
tail -n 4 gpurun_out/r2i_gate_default.log gpurun_out/r2i_gate_item.log; cat gpurun_out/r2i_attn_bench.log
for f in gpurun_out/r2i_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 6; done
for f in r2i_bench_item r2i_bench_default; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; done
