#!/bin/bash
# gpurun batch J (round 2, 1 GPU): ncu source-level capture of the item-level forward attention kernel.
mkdir -p gpurun_out
MMB_ATTN_FWD=item timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd --launch-skip 2 \
   --launch-count 1 -f -o gpurun_out/r2_attn_fwd_item python scripts/ncu_attn_fwd.py > gpurun_out/r2j_ncu.log 2>&1
tail -n 3 gpurun_out/r2j_ncu.log
