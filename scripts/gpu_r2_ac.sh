#!/bin/bash
# gpurun batch AC (round 2, 1 GPU): text tower on a side stream (MMB_TOWER_STREAMS=1) vs the sequential schedule.
mkdir -p gpurun_out
MMB_TOWER_STREAMS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "trainer or full_size or micro" --timeout=500 > gpurun_out/r2ac_test_streams.log 2>&1
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2ac_bench_seq.json 2> gpurun_out/r2ac_bench_seq.err
MMB_TOWER_STREAMS=1 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2ac_bench_streams.json 2> gpurun_out/r2ac_bench_streams.err
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2ac_bench_seq2.json 2> gpurun_out/r2ac_bench_seq2.err
grep -E "^FAILED|passed|failed" gpurun_out/r2ac_test_streams.log | tail -n 4
for f in r2ac_bench_seq r2ac_bench_streams r2ac_bench_seq2; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; tail -n 2 gpurun_out/$f.err | cut -c1-200; done
