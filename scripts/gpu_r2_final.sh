#!/bin/bash
# gpurun final batch (round 2, 1 GPU): what the driver runs at round end, on the committed state — build + smoke, every
# GPU test, the default bench line and the reference arm.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2final_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=900 > gpurun_out/r2final_pytest_gpu.log 2>&1; tail -n 2 gpurun_out/r2final_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2final_bench_reference.json 2> gpurun_out/r2final_bench_reference.err
grep '^{' gpurun_out/r2final_bench.json | head -c 700; echo
grep '^{' gpurun_out/r2final_bench_reference.json | head -c 300; echo
