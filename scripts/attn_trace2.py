"""Debug: per-phase SM-clock trace of the persistent attention-backward kernel (CTA 1, 7th tile).  Needs the trace
build (libmmb200_trace.so: csrc/*.cu compiled with -DMMB_ATTN_TRACE)."""
import ctypes
import os

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
    L.mmb_attention_bwd_tc(vp(qkv.data_ptr()), vp(out.data_ptr()), vp(dout.data_ptr()), vp(lse.data_ptr()),
                           vp(dqkv.data_ptr()), B, S, H, 0, f(0.125), st)
    torch.cuda.synchronize()
t = trace.cpu().view(2, 4, 64)
for kind in (0, 1):
    w, si, ai = t[kind, 0].tolist(), t[kind, 1].tolist(), t[kind, 2].tolist()
    t0 = w[40]
    rel = lambda x: (x - t0) if x else -1  # noqa: E731
    print(f"== persistent {'DKDV' if kind else 'DQ'} (SM clocks since the worker entered the tile)")
    print("  worker: stats_ready", rel(w[41]), "| chunk loop done -> acc wait begin", rel(w[42]), "acc ready", rel(w[43]), "tile done", rel(w[44]))
    for c in range(4):
        print(f"  chunk {c}: worker begin {rel(w[5*c])} scores_ready {rel(w[5*c+1])} tmem_read {rel(w[5*c+2])} ds_buf_free {rel(w[5*c+3])} stored {rel(w[5*c+4])}"
              f" | score-issuer begin {rel(si[4*c])} ring_ok {rel(si[4*c+1])} sdp_free {rel(si[4*c+2])} issued {rel(si[4*c+3])}"
              f" | acc-issuer begin {rel(ai[4*c])} ds_ready {rel(ai[4*c+1])} issued {rel(ai[4*c+2])}")
