#!/bin/bash
# gpurun batch H (round 2, 1 GPU): TMEM-read / MUFU micro-benchmark + full ncu capture of the attention kernels
# (forward tile kernel, fused single-pass backward) at the B/16 shape.
mkdir -p gpurun_out
timeout 120 ./build/tmem_probe > gpurun_out/r2h_tmem_probe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_' --launch-skip 2 --launch-count 2 -f \
   -o gpurun_out/r2_attn_fused python scripts/ncu_kernels.py > gpurun_out/r2h_ncu_attn.log 2>&1
cat gpurun_out/r2h_tmem_probe.log; tail -n 3 gpurun_out/r2h_ncu_attn.log
