#!/bin/bash
# gpurun batch P: issuer / producer stamps of the fused backward
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2p_build.log 2>&1
timeout 200 python scripts/attn_bwd_trace.py > gpurun_out/r2p_bwd_trace.log 2>&1
grep -A 20 "== CTA 0" gpurun_out/r2p_bwd_trace.log | grep -v "worker [1-7]:" | cut -c1-1500
