#!/bin/bash
# gpurun batch M: micro-benchmarks (softmax inner loops, corrected MUFU / FFMA2 / tcgen05.ld probes)
mkdir -p gpurun_out
timeout 120 ./build/softmax_probe > gpurun_out/r2m_softmax_probe.log 2>&1
timeout 120 ./build/tmem_probe > gpurun_out/r2m_tmem_probe.log 2>&1
cat gpurun_out/r2m_softmax_probe.log gpurun_out/r2m_tmem_probe.log
