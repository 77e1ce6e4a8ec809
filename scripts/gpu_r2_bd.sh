#!/bin/bash
# gpurun batch: timings + one `ncu --set full` capture of the kernels added in the re-entry session, then the driver's own
# round-end sequence (build + smoke, pytest -x -m gpu) on the committed state.
mkdir -p gpurun_out
timeout 300 python scripts/ncu_new_kernels.py > gpurun_out/bd_timing.log 2>&1; tail -n 4 gpurun_out/bd_timing.log
NCU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'clip_|attn_gen_bwd' -c 6 \
  -o gpurun_out/r2b_new_kernels -f python scripts/ncu_new_kernels.py > gpurun_out/bd_ncu.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/bd_ncu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/bd_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/bd_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=900 > gpurun_out/bd_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/bd_pytest_gpu.log
