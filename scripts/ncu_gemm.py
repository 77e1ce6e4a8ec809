"""Runs a few representative GEMM launches (for `ncu --set full -k regex:gemm_kernel`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
T = 148 * 128 * 4   # 75776 token rows: 4 full waves of M tiles per n-tile column
def t(*s): return torch.randn(*s, device=dev).bfloat16()
cases = [
    ("qkv_fwd", dict(A=t(T, 768), B=t(2304, 768), epilogue=0, bias=torch.randn(2304, device=dev))),
    ("fc2_fwd", dict(A=t(T, 3072), B=t(768, 3072), epilogue=0, bias=torch.randn(768, device=dev))),
    ("fc1_act", dict(A=t(T, 768), B=t(3072, 768), epilogue=1, bias=torch.randn(3072, device=dev))),
    ("fc1_wgrad", dict(A=t(T, 3072), B=t(T, 768), a_mn=True, b_mn=True, epilogue=3, splits=8)),
]
for rep in range(2):
    for name, kw in cases:
        ops.gemm(**kw)
torch.cuda.synchronize()
print("done")
