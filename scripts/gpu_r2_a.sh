#!/bin/bash
# gpurun batch A (round 2): kernel gates first (a hanging kernel variant is switched off for the rest of the batch),
# then parity tests file by file, kernel micro-benchmarks (A/B inside one box), one bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2a_build.log 2>&1
# gate 1: attention forward (new persistent kernel); gate 2: attention backward (ping-pong kernel)
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2a_gate_attn.log 2>&1
if ! grep -q " passed" gpurun_out/r2a_gate_attn.log || grep -q "failed" gpurun_out/r2a_gate_attn.log; then
  echo "GATE: default attention kernels failed -> trying fwd=tile" >> gpurun_out/r2a_gate_attn.log
  export MMB_ATTN_FWD=tile
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention_fwd_bwd" --timeout=200 > gpurun_out/r2a_gate_attn2.log 2>&1
  if ! grep -q " passed" gpurun_out/r2a_gate_attn2.log || grep -q "failed" gpurun_out/r2a_gate_attn2.log; then
    echo "GATE: still failing -> bwd=colsplit too" >> gpurun_out/r2a_gate_attn2.log
    export MMB_ATTN_BWD=colsplit
  fi
fi
echo "MMB_ATTN_FWD=$MMB_ATTN_FWD MMB_ATTN_BWD=$MMB_ATTN_BWD" > gpurun_out/r2a_variants.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
for f in tests/test_gpu_optim.py tests/test_gpu_parity.py tests/test_gpu_layers.py tests/test_gpu_flava.py tests/test_gpu_coca.py tests/test_gpu_distributed.py; do
  timeout 900 python -m pytest $f -q -m gpu -s --timeout=600 > gpurun_out/r2a_$(basename $f .py).log 2>&1
  echo "rc=$?" >> gpurun_out/r2a_$(basename $f .py).log
done
timeout 300 python scripts/attn_bench.py > gpurun_out/r2a_attn_bench.log 2>&1
MMB_ATTN_BWD=colsplit MMB_ATTN_FWD=tile timeout 300 python scripts/attn_bench.py >> gpurun_out/r2a_attn_bench.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
cat gpurun_out/r2a_variants.txt; tail -n 3 gpurun_out/r2a_gate_attn.log; tail -c 400 gpurun_out/r2a_smoke.log
for f in gpurun_out/r2a_test_*.log; do echo "== $f"; grep -E "^FAILED|passed|failed" $f | tail -n 12; done
grep -v watchdog gpurun_out/r2a_attn_bench.log | tail -n 12; head -c 3000 gpurun_out/r2a_bench.json; tail -n 5 gpurun_out/r2a_bench.err
