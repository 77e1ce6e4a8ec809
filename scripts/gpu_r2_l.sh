#!/bin/bash
# gpurun batch L (round 2, 1 GPU): phase trace of the item forward kernel; L2 prefetch A/B in the fused backward.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2l_build.log 2>&1
timeout 200 python scripts/attn_item_trace.py > gpurun_out/r2l_item_trace.log 2>&1
MMB_ATTN_ITEM_STAGGER=0 timeout 200 python scripts/attn_item_trace.py > gpurun_out/r2l_item_trace_lockstep.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2l_gate_default.log 2>&1
( echo "== bwd fused + L2 prefetch"; timeout 300 python scripts/attn_bench.py
  echo "== bwd fused, no prefetch"; MMB_ATTN_L2PF=0 timeout 300 python scripts/attn_bench.py ) 2>&1 | grep -v watchdog > gpurun_out/r2l_attn_bench.log
cat gpurun_out/r2l_item_trace.log; echo; cat gpurun_out/r2l_item_trace_lockstep.log | head -12; tail -n 3 gpurun_out/r2l_gate_default.log; cat gpurun_out/r2l_attn_bench.log
