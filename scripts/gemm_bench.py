"""Times the GEMM kernel alone at the headline shapes of the CLIP ViT-B/16 step (B = 1024: M = 201 728 image tokens),
for the default dispatch and for the 16-epilogue-warp activation variant, next to cuBLAS on the same shapes.
CUDA events, 10 iterations after warm-up.  Run under gpurun; appends to gpurun_out/r2_gemm_bench.log."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
lines = []


def log(s):
    print(s, flush=True)
    lines.append(s)


def timeit(fn, iters=10):
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


def bench(name, M, N, K, a_mn, b_mn, epi, splits=1, colsum=False, bias=True):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    aux = torch.randn(M, N, device=dev).bfloat16() if epi == 2 else None
    bv = torch.randn(N, device=dev) if (bias and epi != 2) else None
    D0 = torch.empty((M, N), device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16)
    D1 = torch.empty((M, N), device=dev, dtype=torch.bfloat16) if epi == 1 else None
    cs = torch.zeros(N, device=dev) if colsum else None
    ms = timeit(lambda: ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, epilogue=epi, out=D0, out2=D1, bias=bv, aux=aux, splits=splits,
                                 accumulate=(epi == 3), colsum=cs))
    A2 = A.t() if a_mn else A
    B2 = B.t() if b_mn else B
    Dc = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    ms2 = timeit(lambda: torch.matmul(A2, B2.t(), out=Dc))
    fl = 2.0 * M * N * K
    log(f"{name:12s} M{M} N{N} K{K} epi={epi} colsum={int(colsum)}: {ms:.3f} ms {fl / ms / 1e9:7.1f} TF/s | cuBLAS (plain bf16 out) "
        f"{ms2:.3f} ms {fl / ms2 / 1e9:7.1f} TF/s")


def suite(tag):
    log(f"---- {tag} ----")
    T, Tt = 1024 * 197, 1024 * 77
    bench("qkv_fwd", T, 2304, 768, 0, 0, 0)
    bench("out_fwd", T, 768, 768, 0, 0, 0)
    bench("fc1_fwd+act", T, 3072, 768, 0, 0, 1)
    bench("fc2_fwd", T, 768, 3072, 0, 0, 0)
    bench("fc2_dgrad*act'", T, 3072, 768, 0, 1, 2, colsum=True)
    bench("fc1_dgrad", T, 768, 3072, 0, 1, 0, bias=False)
    bench("qkv_dgrad", T, 768, 2304, 0, 1, 0, bias=False)
    bench("fc1_wgrad", 3072, 768, T, 1, 1, 3, splits=ops.wgrad_splits(3072, 768, T), bias=False)
    bench("qkv_wgrad", 2304, 768, T, 1, 1, 3, splits=ops.wgrad_splits(2304, 768, T), bias=False)
    bench("txt fc1+act", Tt, 2048, 512, 0, 0, 1)
    bench("txt fc2_dgrad", Tt, 2048, 512, 0, 1, 2, colsum=True)


log(f"device: {torch.cuda.get_device_name(0)}")
suite("default dispatch (CTA pairs, 8 epilogue warps)")
_lib.lib().mmb_gemm_set_mode(-1, 16)
suite("activation epilogues with 16 epilogue warps")
_lib.lib().mmb_gemm_set_mode(-1, 0)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/r2_gemm_bench.log", "a") as f:
    f.write("\n".join(lines) + "\n")
