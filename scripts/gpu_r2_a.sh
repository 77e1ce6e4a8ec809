#!/bin/bash
# gpurun batch A (round 2): parity tests file by file (a CUDA fault in one file cannot poison the others), kernel
# micro-benchmarks (A/B inside one box), one bench line, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
for f in tests/test_gpu_optim.py tests/test_gpu_parity.py tests/test_gpu_flava.py tests/test_gpu_coca.py tests/test_gpu_distributed.py; do
  timeout 900 python -m pytest $f -q -m gpu -s --timeout=600 > gpurun_out/r2a_$(basename $f .py).log 2>&1
  echo "rc=$?" >> gpurun_out/r2a_$(basename $f .py).log
done
timeout 300 python scripts/attn_bench.py > gpurun_out/r2a_attn_bench.log 2>&1
MMB_ATTN_BWD=colsplit timeout 300 python scripts/attn_bench.py >> gpurun_out/r2a_attn_bench.log 2>&1
rm -f gpurun_out/r2_gemm_bench.log; timeout 600 python scripts/gemm_bench.py > gpurun_out/r2a_gemm_bench.out 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_smoke.log; for f in gpurun_out/r2a_test_*.log; do echo "== $f"; tail -n 6 $f; done
cat gpurun_out/r2a_attn_bench.log; tail -n 30 gpurun_out/r2a_gemm_bench.out; head -c 1500 gpurun_out/r2a_bench.json
