#!/bin/bash
# gpurun final batch of the re-entry session (round 2, 1 GPU): what the driver runs at round end, on the committed state —
# build + smoke, EVERY GPU test (no -x here: all failures in one pass), the new probes, the default bench line and the
# reference arm.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/f2_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/f2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/f2_pytest_gpu.log
timeout 300 python scripts/clip_transform_probe.py > gpurun_out/f2_transform_probe.log 2>&1; tail -n 3 gpurun_out/f2_transform_probe.log
timeout 400 python scripts/flava_train_probe.py > gpurun_out/f2_flava_train_probe.log 2>&1; tail -n 2 gpurun_out/f2_flava_train_probe.log
timeout 900 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; grep '^{' gpurun_out/f2_bench.json | head -c 600; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f2_bench_reference.json 2> gpurun_out/f2_bench_reference.err; grep '^{' gpurun_out/f2_bench_reference.json | head -c 300; echo
