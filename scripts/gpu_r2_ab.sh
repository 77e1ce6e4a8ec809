#!/bin/bash
# gpurun batch AB: FLAVA tests after the restructuring of the loss module
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flava_pretraining.py tests/test_gpu_flava.py -q -m gpu --timeout=500 > gpurun_out/r2ab_test.log 2>&1
tail -n 15 gpurun_out/r2ab_test.log | cut -c1-300
