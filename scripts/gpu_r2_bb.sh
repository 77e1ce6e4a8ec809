#!/bin/bash
# gpurun batch (round 2, re-entry): FLAVA heads backward + end-to-end pre-training step, then the timing probe.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flava_train.py -q --timeout=400 -s > gpurun_out/bb_flava_train.log 2>&1; echo "flava_train rc=$?"; tail -n 4 gpurun_out/bb_flava_train.log
timeout 600 python -m pytest tests/test_gpu_flava_pretraining.py -q --timeout=400 > gpurun_out/bb_flava_pre.log 2>&1; echo "flava_pre rc=$?"; tail -n 3 gpurun_out/bb_flava_pre.log
timeout 400 python -m pytest tests/test_gpu_coca_train.py tests/test_gpu_coca.py -q --timeout=300 -s > gpurun_out/bb_coca_train.log 2>&1; echo "coca_train rc=$?"; tail -n 3 gpurun_out/bb_coca_train.log
timeout 300 python -m pytest tests/test_gpu_clip_transform.py -q --timeout=200 > gpurun_out/bb_clip_transform.log 2>&1; echo "clip_transform rc=$?"; tail -n 3 gpurun_out/bb_clip_transform.log
timeout 300 python scripts/clip_transform_probe.py > gpurun_out/bb_transform_probe.log 2>&1; tail -n 3 gpurun_out/bb_transform_probe.log
timeout 400 python scripts/flava_train_probe.py > gpurun_out/bb_probe.log 2>&1; echo "probe rc=$?"; tail -n 3 gpurun_out/bb_probe.log
