"""HBM bandwidth by access mix (CUDA events, 4 GiB buffers, well above the 126 MB L2): write-only (memset), read-only
(sum), copy (read + write).  Says which floor a store-heavy epilogue (FC1 + QuickGELU writes 2.4 GB, reads 0.3 GB) is up
against."""
import torch

dev = torch.device("cuda:0")
n = 1 << 30   # fp32 elements: 4 GiB
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


gb = n * 4 / 1e9
print(f"write-only (fill_):   {gb / t(lambda: a.fill_(1.0)):7.0f} GB/s")
print(f"write-only (zero_):   {gb / t(lambda: a.zero_()):7.0f} GB/s")
print(f"read-only  (sum):     {gb / t(lambda: a.sum()):7.0f} GB/s")
print(f"copy (read + write):  {2 * gb / t(lambda: b.copy_(a)):7.0f} GB/s total")
ab = a.view(torch.bfloat16)
print(f"read 1 : write 8 (bf16 -> 8 x ...): skipped; read 4 B + write 2 B (float -> bf16 cast): "
      f"{(gb + gb / 2) / t(lambda: torch.empty(n, device=dev, dtype=torch.bfloat16).copy_(a)) :7.0f} GB/s total (incl. allocation)")
