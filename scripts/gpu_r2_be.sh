#!/bin/bash
# gpurun batch: the register-cached single-sweep path of the general attention backward (query kernel) — parity tests of
# everything that uses it, then its timing against the first version (profiles/r2b_new_kernels_timing_first.log).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_coca_train.py -q --timeout=400 > gpurun_out/be_coca_train.log 2>&1; echo "coca_train rc=$?"; tail -n 3 gpurun_out/be_coca_train.log
timeout 300 python scripts/ncu_new_kernels.py > gpurun_out/be_timing.log 2>&1; tail -n 4 gpurun_out/be_timing.log
