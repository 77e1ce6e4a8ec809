#!/bin/bash
# gpurun batch V (round 2, 1 GPU): fused backward chunk loop split into an exp phase (under the dP^T load, before the
# buffer waits) and a store phase.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2v_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2v_gate_default.log 2>&1
timeout 300 python scripts/attn_bench.py 2>&1 | grep -v watchdog > gpurun_out/r2v_attn_bench.log
timeout 200 python scripts/attn_bwd_trace.py > gpurun_out/r2v_bwd_trace.log 2>&1
if grep -q " passed" gpurun_out/r2v_gate_default.log && ! grep -q "failed" gpurun_out/r2v_gate_default.log; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flava.py tests/test_gpu_distributed.py -q -m gpu --timeout=600 > gpurun_out/r2v_test_gpu.log 2>&1
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
fi
tail -n 3 gpurun_out/r2v_gate_default.log; cat gpurun_out/r2v_attn_bench.log
grep -A 2 "== CTA 0" gpurun_out/r2v_bwd_trace.log | cut -c1-500
for f in gpurun_out/r2v_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 6; done
grep '^{' gpurun_out/r2v_bench.json | head -c 330; echo
