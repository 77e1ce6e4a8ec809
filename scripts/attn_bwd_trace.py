"""Per-phase SM-clock breakdown of the fused attention backward kernel's worker warps (CTAs 0-1), B/16 shape."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, H = 1024, 197, 12
d = H * 64
qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).bfloat16()
out = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H * S, device=dev)
dout = (torch.randn(B * S, d, device=dev) * 0.5).bfloat16()
dqkv = torch.empty_like(qkv)
ops.attention_fwd(qkv, out, lse, B, S, H, False, 0.125)
for _ in range(3):
    ops.attention_bwd(qkv, out, dout, lse, dqkv, B, S, H, False, 0.125)
torch.cuda.synchronize()
trace = torch.zeros(4 * 12 * 64, dtype=torch.int64, device=dev)
L = _lib.lib()
L.mmb_debug_attn_item_trace.argtypes = [ctypes.c_void_p]
L.mmb_debug_attn_item_trace(ctypes.c_void_p(trace.data_ptr()))
ops.attention_bwd(qkv, out, dout, lse, dqkv, B, S, H, False, 0.125)
torch.cuda.synchronize()
L.mmb_debug_attn_item_trace(ctypes.c_void_p(0))
t = trace.cpu().view(4, 12, 64)
names = ["wait stats", "wait S/dP", "tmem ld", "wait bufs", "compute", "wait dK/dV", "dK/dV epi", "wait dQ", "dQ epi"]
for cta in range(2):
    print(f"== CTA {cta}")
    for w in range(8):
        a = t[cta, w]
        n = max(1, int(a[9]))
        print(f"  worker {w}: items {int(a[9])}  per item: " + " | ".join(f"{nm} {int(a[i]) // n}" for i, nm in enumerate(names)) +
              f" | first-tile epi {int(a[10]) // n} | sum {sum(int(a[i]) for i in range(9)) // n}")
    base = int(t[cta, 0, 16])
    st = [(int(t[cta, 0, 16 + 2 * i]) - base, int(t[cta, 0, 17 + 2 * i]) - base) for i in range(20)]
    print(f"  worker 0: (S/dP ready, dS done) of chunks 0-19, clocks: {st}")
    print(f"  producer: ring load issue of chunks 0-19: {[int(t[cta, 8, i]) - base for i in range(20)]}")
    print(f"  score issuer: (ring_full seen, issue) of chunks 0-19: {[(int(t[cta, 9, 2 * i]) - base, int(t[cta, 9, 2 * i + 1]) - base) for i in range(20)]}")
    print(f"  acc issuer: (dS seen, MMAs issued) of chunks 0-19: {[(int(t[cta, 10, 2 * i]) - base, int(t[cta, 10, 2 * i + 1]) - base) for i in range(20)]}")
    print(f"  worker 0: dK/dV final seen, key tiles 0-5: {[int(t[cta, 11, i]) - base for i in range(6)]}")
