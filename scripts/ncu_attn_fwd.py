"""Launches the attention forward at the B/16 shape (B = 1024, S = 197, H = 12) a few times for an ncu capture:
    MMB_ATTN_FWD=item ncu --set full --clock-control none --import-source on -k regex:attn_fwd --launch-skip 2 \
        --launch-count 1 -o gpurun_out/r2_attn_fwd_item python scripts/ncu_attn_fwd.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, H = 1024, 197, 12
d = H * 64
qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).bfloat16()
out = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H * S, device=dev)
for _ in range(3):
    ops.attention_fwd(qkv, out, lse, B, S, H, False, 0.125)
torch.cuda.synchronize()
print("done")
