#!/bin/bash
# gpurun batch Q (round 2, 1 GPU): fused backward with TMA-store epilogues, early acc_full commit, hoisted descriptors.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2q_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2q_gate_default.log 2>&1
timeout 300 python scripts/attn_bench.py 2>&1 | grep -v watchdog > gpurun_out/r2q_attn_bench.log
timeout 200 python scripts/attn_bwd_trace.py > gpurun_out/r2q_bwd_trace.log 2>&1
if grep -q " passed" gpurun_out/r2q_gate_default.log && ! grep -q "failed" gpurun_out/r2q_gate_default.log; then
  for f in parity flava coca distributed; do
    timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout=600 > gpurun_out/r2q_test_gpu_$f.log 2>&1
  done
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
fi
tail -n 3 gpurun_out/r2q_gate_default.log; cat gpurun_out/r2q_attn_bench.log
grep -A 3 "== CTA 0" gpurun_out/r2q_bwd_trace.log | cut -c1-500
grep -A 16 "== CTA 0" gpurun_out/r2q_bwd_trace.log | grep -v "worker [0-7]: items" | cut -c1-1300
for f in gpurun_out/r2q_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 6; done
grep '^{' gpurun_out/r2q_bench.json | head -c 330; echo
