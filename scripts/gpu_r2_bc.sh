#!/bin/bash
# gpurun batch: re-run of the three tests fixed after the final batch (test-side fixes: forward tolerance of the FLAVA
# base-width gradient test; the CoCa oracle runs on the CPU).
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_coca_train.py" "tests/test_gpu_flava_train.py::test_flava_base_width_gradients_against_fp32_oracle" -q --timeout=400 -s > gpurun_out/bc_fixed.log 2>&1; echo "rc=$?"; tail -n 4 gpurun_out/bc_fixed.log
