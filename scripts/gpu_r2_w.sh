#!/bin/bash
# gpurun batch W (round 2, 1 GPU): packed f32x2 QuickGELU epilogues (FC1 + act, FC2-dgrad x act') in the GEMM.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; tail -n 1 gpurun_out/r2w_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=500 > gpurun_out/r2w_test_gpu_parity.log 2>&1
timeout 600 python scripts/gemm_bench.py > gpurun_out/r2w_gemm_bench.log 2>&1
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_flava.py tests/test_gpu_coca.py tests/test_gpu_distributed.py tests/test_gpu_optim.py -q -m gpu --timeout=500 > gpurun_out/r2w_test_gpu_rest.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
MMB_FUSE_COLSUM_GEMM=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2w_bench_nofuse.json 2> gpurun_out/r2w_bench_nofuse.err
grep -E "^FAILED|passed|failed" gpurun_out/r2w_test_gpu_parity.log gpurun_out/r2w_test_gpu_rest.log | tail -n 8
tail -n 12 gpurun_out/r2w_gemm_bench.log
for f in r2w_bench r2w_bench_nofuse; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; done
