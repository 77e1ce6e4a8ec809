"""Runs the attention kernels once at the headline shape (for `ncu --set full -k regex:attn`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, H = 256, 197, 12
d = H * 64
qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).bfloat16()
out = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H * S, device=dev)
dout = (torch.randn(B * S, d, device=dev) * 0.5).bfloat16()
dqkv = torch.empty_like(qkv)
for _ in range(2):
    ops.attention_fwd(qkv, out, lse, B, S, H, False, 0.125)
    ops.attention_bwd(qkv, out, dout, lse, dqkv, B, S, H, False, 0.125)
torch.cuda.synchronize()
print("done")
