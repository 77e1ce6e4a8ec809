"""Times the attention kernels alone at the headline shapes (CUDA events, 20 iterations after warm-up).
MMB_ATTN_BWD_PERSIST=0|1 selects the backward variant (read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(B, S, H, causal, iters=20):
    d = H * 64
    qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).bfloat16()
    out = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B * H * S, device=dev)
    dout = (torch.randn(B * S, d, device=dev) * 0.5).bfloat16()
    dqkv = torch.empty_like(qkv)

    def t(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    f = t(lambda: ops.attention_fwd(qkv, out, lse, B, S, H, causal, 0.125))
    b = t(lambda: ops.attention_bwd(qkv, out, dout, lse, dqkv, B, S, H, causal, 0.125))
    fl = 4.0 * S * S * 64 * H * B * (0.5 if causal else 1.0)
    print(f"B={B} S={S} H={H} causal={causal} persist={os.environ.get('MMB_ATTN_BWD_PERSIST', '1')}: fwd {f:.3f} ms "
          f"({fl / f / 1e9:.0f} TF/s)  bwd {b:.3f} ms ({2.5 * fl / b / 1e9:.0f} TF/s)", flush=True)


run(1024, 197, 12, False)
run(1024, 77, 8, True)
run(256, 197, 12, False)
