"""Launches ONE instance of every hot kernel at the shapes of the benchmarked step (CLIP ViT-B/16, B = 1024 per GPU:
M = 201 728 image tokens) so that `ncu --set full` can capture the exact launches bench.py times:

    ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn_|add_ln_fwd|ln_bwd' \
        --launch-skip 14 -o gpurun_out/r2_kernels python scripts/ncu_kernels.py

Every kernel is launched twice (the first, un-profiled pass — skipped with --launch-skip — warms instruction caches
and creates the tensor maps); scripts/ncu_summarize.py turns the report into profiles/r2_ncu_kernels.{csv,json}."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, H, d, ff = 1024, 197, 12, 768, 3072
M = B * S
bf, f32 = torch.bfloat16, torch.float32
x = (torch.randn(M, d, device=dev) * 0.5).to(bf)
w_qkv = (torch.randn(3 * d, d, device=dev) * 0.03).to(bf)
w_fc1 = (torch.randn(ff, d, device=dev) * 0.03).to(bf)
w_fc2 = (torch.randn(d, ff, device=dev) * 0.03).to(bf)
b_qkv, b_fc1 = torch.randn(3 * d, device=dev), torch.randn(ff, device=dev)
qkv = torch.empty(M, 3 * d, device=dev, dtype=bf)
pre, hact = torch.empty(M, ff, device=dev, dtype=bf), torch.empty(M, ff, device=dev, dtype=bf)
gb = (torch.randn(M, d, device=dev) * 0.1).to(bf)
dpre = torch.empty(M, ff, device=dev, dtype=bf)
cs = torch.zeros(ff, device=dev)
dw1 = torch.zeros(ff, d, device=dev)
o = torch.empty(M, d, device=dev, dtype=bf)
lse = torch.empty(B * H * S, device=dev)
dqkv = torch.empty_like(qkv)
xs, xo = torch.randn(M, d, device=dev), torch.empty(M, d, device=dev)
ln = torch.empty(M, d, device=dev, dtype=bf)
g, bta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
G, dg, db, gs = torch.randn(M, d, device=dev) * 0.1, torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
Gb = torch.empty(M, d, device=dev, dtype=bf)


def once():
    ops.gemm(x, w_qkv, bias=b_qkv, out=qkv)                                                     # 1 gemm<K,K,bf16> QKV fwd
    ops.gemm(x, w_fc1, bias=b_fc1, epilogue=ops.EPI_BF16_ACT, out=pre, out2=hact)               # 2 FC1 + QuickGELU
    ops.gemm(gb, w_fc2, b_mn=True, epilogue=ops.EPI_BF16_DACT, aux=pre, out=dpre, colsum=cs)    # 3 FC2 dgrad x act' + colsum
    ops.gemm(dpre, x, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=dw1, splits=ops.wgrad_splits(ff, d, M),
             accumulate=True)                                                                   # 4 FC1 wgrad (split-K)
    ops.attention_fwd(qkv, o, lse, B, S, H, False, 0.125)                                       # 5 attn fwd
    ops.attention_bwd(qkv, o, gb, lse, dqkv, B, S, H, False, 0.125)                             # 6 attn bwd (fused: one kernel)
    ops.add_layernorm_fwd(xs, gb, xo, ln, None, g, bta, mean, rstd, M, d, 1e-5)                 # 8 add + LayerNorm fwd
    ops.layernorm_bwd(xo, gb, None, mean, rstd, g, G, G, Gb, dg, db, M, d, gsum=gs)             # 9 LayerNorm bwd


once()
torch.cuda.synchronize()
once()
torch.cuda.synchronize()
print("done")
