#!/bin/bash
# gpurun batch Y (round 2, 1 GPU): forward max pass with two loads per wait; FLAVA pre-training loss timing probe.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2y_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2y_gate.log 2>&1
timeout 300 python scripts/attn_bench.py 2>&1 | grep -v watchdog > gpurun_out/r2y_attn_bench.log
timeout 200 python scripts/attn_item_trace.py > gpurun_out/r2y_item_trace.log 2>&1
timeout 300 python scripts/flava_pretraining_probe.py > gpurun_out/r2y_flava_pretraining_probe.log 2>&1
tail -n 2 gpurun_out/r2y_gate.log; cat gpurun_out/r2y_attn_bench.log; head -n 4 gpurun_out/r2y_item_trace.log | cut -c1-300; tail -n 2 gpurun_out/r2y_flava_pretraining_probe.log
