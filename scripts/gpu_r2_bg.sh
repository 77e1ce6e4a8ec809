#!/bin/bash
# gpurun batch: shared-memory tiled general attention backward — parity of everything that uses it, kernel timings and the
# CoCa train-step probe (A/B against MMB_ATTN_GEN_BWD=rows), then the driver's round-end sequence on the committed state.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_coca_train.py -q --timeout=300 > gpurun_out/bg_coca_train.log 2>&1; echo "coca_train rc=$?"; tail -n 3 gpurun_out/bg_coca_train.log
timeout 200 python scripts/ncu_new_kernels.py > gpurun_out/bg_timing.log 2>&1; tail -n 2 gpurun_out/bg_timing.log
timeout 200 python scripts/coca_train_probe.py > gpurun_out/bg_probe_tiled.log 2>&1; tail -n 1 gpurun_out/bg_probe_tiled.log
MMB_ATTN_GEN_BWD=rows timeout 200 python scripts/coca_train_probe.py > gpurun_out/bg_probe_rows.log 2>&1; tail -n 1 gpurun_out/bg_probe_rows.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/bg_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/bg_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout=600 > gpurun_out/bg_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/bg_pytest_gpu.log
