#!/bin/bash
# gpurun batch AD: attention parity incl. causal two-tile sequences
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=500 > gpurun_out/r2ad_gate.log 2>&1
MMB_ATTN_BWD=colsplit timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=500 > gpurun_out/r2ad_gate_colsplit.log 2>&1
tail -n 12 gpurun_out/r2ad_gate.log | cut -c1-300; tail -n 3 gpurun_out/r2ad_gate_colsplit.log
