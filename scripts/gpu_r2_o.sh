#!/bin/bash
# gpurun batch O (round 2, 1 GPU): elect_one / uniform-warp issue paths (no ELECT+R2UR waterfall per tcgen05.mma / TMA)
# in every attention kernel and the GEMM; item forward kernel with TMA-store epilogue.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2o_build_smoke.log 2>&1
tail -n 2 gpurun_out/r2o_build_smoke.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd or gemm" --timeout=200 > gpurun_out/r2o_gate_default.log 2>&1
MMB_ATTN_FWD=item timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2o_gate_item.log 2>&1
MMB_ATTN_BWD=colsplit MMB_ATTN_FWD=pp timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2o_gate_colsplit_pp.log 2>&1
MMB_ATTN_BWD=pp timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2o_gate_bwdpp.log 2>&1
( echo "== fwd tile (default), bwd fused"; timeout 300 python scripts/attn_bench.py
  echo "== fwd item"; MMB_ATTN_FWD=item timeout 300 python scripts/attn_bench.py
  echo "== fwd pp"; MMB_ATTN_FWD=pp timeout 300 python scripts/attn_bench.py ) 2>&1 | grep -v watchdog > gpurun_out/r2o_attn_bench.log
timeout 200 python scripts/attn_item_trace.py > gpurun_out/r2o_item_trace.log 2>&1
timeout 200 python scripts/attn_bwd_trace.py > gpurun_out/r2o_bwd_trace.log 2>&1
timeout 600 python scripts/gemm_bench.py > gpurun_out/r2o_gemm_bench.log 2>&1
for f in parity optim layers flava coca distributed; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --timeout=600 > gpurun_out/r2o_test_gpu_$f.log 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err
MMB_ATTN_FWD=item timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2o_bench_item.json 2> gpurun_out/r2o_bench_item.err
tail -n 3 gpurun_out/r2o_gate_*.log; cat gpurun_out/r2o_attn_bench.log
head -n 12 gpurun_out/r2o_item_trace.log | cut -c1-400; head -n 10 gpurun_out/r2o_bwd_trace.log | cut -c1-400
tail -n 12 gpurun_out/r2o_gemm_bench.log
for f in gpurun_out/r2o_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 6; done
for f in r2o_bench_default r2o_bench_item; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; done
