#!/bin/bash
# gpurun batch AF: tcgen05.mma with A in tensor memory — layout check and cost per 128x64x16 MMA
mkdir -p gpurun_out
timeout 60 ./build/mma_ts_probe > gpurun_out/r2af_mma_ts_probe.log 2>&1; echo "rc=$?"
cat gpurun_out/r2af_mma_ts_probe.log
