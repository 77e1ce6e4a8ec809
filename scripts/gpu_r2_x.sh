#!/bin/bash
# gpurun batch X (round 2, 1 GPU): FLAVA pre-training losses on the CUDA path
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2x_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_flava_pretraining.py -q -m gpu --timeout=500 -x > gpurun_out/r2x_test_flava_pretraining.log 2>&1
tail -n 40 gpurun_out/r2x_test_flava_pretraining.log | cut -c1-400
