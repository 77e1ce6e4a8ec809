"""Debug: per-phase SM-clock trace of one attention-backward CTA per (kernel, tile).  Needs the trace build
(libmmb200_trace.so: every csrc/*.cu compiled with -DMMB_ATTN_TRACE).  Usage: MMB_ATTN_NG=2|4 python scripts/attn_trace.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = ctypes.CDLL(os.path.join(ROOT, "multimodal_b200", "libmmb200_trace.so"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, H = 256, 197, 12
d = H * 64
qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).bfloat16()
out = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B * H * S, device=dev)
dout = (torch.randn(B * S, d, device=dev) * 0.5).bfloat16()
dqkv = torch.empty_like(qkv)
vp = ctypes.c_void_p
st = vp(torch.cuda.current_stream().cuda_stream)
f = ctypes.c_float
L.mmb_attention_fwd_tc(vp(qkv.data_ptr()), vp(out.data_ptr()), vp(lse.data_ptr()), B, S, H, 0, f(0.125), st)
trace = torch.zeros(8 * 64, dtype=torch.int64, device=dev)
for it in range(3):
    trace.zero_()
    torch.cuda.synchronize()
    L.mmb_debug_attn_trace(vp(trace.data_ptr()))
    rc = L.mmb_attention_bwd_tc(vp(qkv.data_ptr()), vp(out.data_ptr()), vp(dout.data_ptr()), vp(lse.data_ptr()),
                                vp(dqkv.data_ptr()), B, S, H, 0, f(0.125), st)
    torch.cuda.synchronize()
t = trace.cpu().view(2, 2, 2, 64)   # kind, tile, role, slot
for kind in (0, 1):
    for tile in (0, 1):
        w, i = t[kind, tile, 0].tolist(), t[kind, tile, 1].tolist()
        t0 = w[0]
        rel = lambda x: (x - t0) if x else -1  # noqa: E731
        print(f"== kernel {'DKDV' if kind else 'DQ'} tile {tile} (SM clocks since CTA entry; NG={os.environ.get('MMB_ATTN_NG', '4')})")
        print("  worker: setup_done", rel(w[1]), "prologue_done", rel(w[2]), "all_mma_done", rel(w[3]), "stores_done", rel(w[4]),
              "exit", rel(w[5]))
        for c in range(4):
            print(f"  chunk {c}: worker scores_ready {rel(w[8+4*c])} tmem_read {rel(w[9+4*c])} bufs_free {rel(w[10+4*c])} "
                  f"stored {rel(w[11+4*c])} | issuer next_scores_begin {rel(i[8+4*c])} issued {rel(i[9+4*c])} "
                  f"ds_seen {rel(i[10+4*c])} acc_issued {rel(i[11+4*c])}")
        print("  issuer: operands_landed", rel(i[0]), "scores0_issued", rel(i[1]))
