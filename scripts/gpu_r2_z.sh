#!/bin/bash
# gpurun batch Z (round 2, 1 GPU): item forward kernel with the O staging in the last P atom (store wait inside pass 2).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2z_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2z_gate.log 2>&1
timeout 300 python scripts/attn_bench.py 2>&1 | grep -v watchdog > gpurun_out/r2z_attn_bench.log
timeout 200 python scripts/attn_item_trace.py > gpurun_out/r2z_item_trace.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flava.py tests/test_gpu_layers.py -q -m gpu --timeout=600 > gpurun_out/r2z_test_gpu.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -n 2 gpurun_out/r2z_gate.log; cat gpurun_out/r2z_attn_bench.log; head -n 4 gpurun_out/r2z_item_trace.log | cut -c1-300
grep -E "^FAILED|passed|failed" gpurun_out/r2z_test_gpu.log | tail -n 4
grep '^{' gpurun_out/r2z_bench.json | head -c 330; echo
