"""Per-phase SM-clock breakdown of the item-level forward attention kernel (MMB_ATTN_FWD=item), CTAs 0-3:
softmax warps: wait S | pass 1 (max) | pass 2 (exp, P) | wait O | epilogue, plus the start stamps of the first items
(is the anti-phase start of the two query tiles kept?)."""
import ctypes
import os
import sys

import torch

os.environ.setdefault("MMB_ATTN_FWD", "item")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _lib, ops  # noqa: E402

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
trace = torch.zeros(4 * 12 * 64, dtype=torch.int64, device=dev)
L = _lib.lib()
L.mmb_debug_attn_item_trace.argtypes = [ctypes.c_void_p]
L.mmb_debug_attn_item_trace(ctypes.c_void_p(trace.data_ptr()))
ops.attention_fwd(qkv, out, lse, B, S, H, False, 0.125)
torch.cuda.synchronize()
L.mmb_debug_attn_item_trace(ctypes.c_void_p(0))
t = trace.cpu().view(4, 12, 64)
for cta in range(2):
    print(f"== CTA {cta}")
    for w in range(8):
        a = t[cta, w]
        n = max(1, int(a[5]))
        print(f"  softmax warp {w} (tile {w // 4}): items {int(a[5])}  per item: wait S {int(a[0]) // n:6d} | pass1 {int(a[1]) // n:6d} | "
              f"pass2 {int(a[2]) // n:6d} | wait O {int(a[3]) // n:6d} | epilogue {int(a[4]) // n:6d} | sum "
              f"{sum(int(a[i]) for i in range(5)) // n}")
    base = int(t[cta, 0, 8])
    for w in (0, 4):
        st = [(int(t[cta, w, 8 + 2 * i]) - base, int(t[cta, w, 9 + 2 * i]) - base) for i in range(10)]
        print(f"  warp {w}: (S ready, P done) of items 0-9, clocks since the first: {st}")
    for w in (9, 10):
        st = [(int(t[cta, w, 2 * i]) - base, int(t[cta, w, 2 * i + 1]) - base) for i in range(10)]
        print(f"  MMA warp {w}: (QK issue, PV issue) of items 0-9: {st}")
