#!/bin/bash
# gpurun batch (round 2, re-entry): FLAVA training path — new backward kernels, gradient parity against the fp32 oracle,
# and the tests that touch what engine.py / attention_tc.cu changed (CLIP step gradients, attention, FLAVA inference).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flava_train.py -q --timeout=400 -x -s > gpurun_out/ba_flava_train.log 2>&1; echo "flava_train rc=$?"; tail -n 3 gpurun_out/ba_flava_train.log
timeout 900 python -m pytest tests/test_gpu_flava.py tests/test_gpu_flava_pretraining.py tests/test_gpu_layers.py -q --timeout=400 > gpurun_out/ba_flava_infer.log 2>&1; echo "flava infer rc=$?"; tail -n 3 gpurun_out/ba_flava_infer.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout=600 -k "attention or full_size or trainer or micro or golden or l14" > gpurun_out/ba_parity.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/ba_parity.log
