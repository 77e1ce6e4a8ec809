"""GPU probe: validates mmb_gemm_bf16 (all operand-major / epilogue combinations) against torch fp32 matmul
and times the headline shapes.  Run under gpurun; writes gpurun_out/gemm_probe.log."""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
out_lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out_lines.append(s)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def gemm(A, a_mn, B, b_mn, M, N, K, epi, act=0, alpha=1.0, bias=None, aux=None, splits=1, accumulate=0, D0=None, D1=None):
    if epi == 3:
        D0 = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32) if D0 is None else D0
        D1 = None
    else:
        D0 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16) if D0 is None else D0
        if epi == 1 and D1 is None:
            D1 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.mmb_gemm_bf16(ptr(A), A.stride(0), a_mn, ptr(B), B.stride(0), b_mn, ptr(D0), D0.stride(0), ptr(D1),
                         D1.stride(0) if D1 is not None else 0, M, N, K, epi, act, alpha, ptr(bias), ptr(aux),
                         aux.stride(0) if aux is not None else 0, splits, accumulate, ctypes.c_void_p(0), ctypes.c_void_p(st))
    if rc != 0:
        raise RuntimeError(f"mmb_gemm_bf16 rc={rc}")
    return D0, D1


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def check(name, got, ref, tol):
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    ok = bool(err <= tol * max(scale, 1e-6)) and bool(torch.isfinite(got).all())
    log(f"{'PASS' if ok else 'FAIL'} {name}: max_abs_err={err:.4e} ref_max={scale:.4e}")
    if not ok:
        bad = ((got - ref).abs() > tol * scale) | ~torch.isfinite(got)
        idx = bad.nonzero()
        log(f"   bad={bad.sum().item()}/{bad.numel()} first={idx[:6].tolist()} rows_bad={bad.any(1).sum().item()} cols_bad={bad.any(0).sum().item()}")
        log("   got[0,:8]=", got[0, :8].tolist(), " ref[0,:8]=", ref[0, :8].tolist())
    return ok


def run_case(M, N, K, a_mn, b_mn, epi, splits=1):
    A2 = torch.randn(M, K, device=dev).bfloat16()  # logical A [M,K]
    B2 = torch.randn(N, K, device=dev).bfloat16()  # logical B [N,K]
    A = A2.t().contiguous() if a_mn else A2
    B = B2.t().contiguous() if b_mn else B2
    bias = torch.randn(N, device=dev)
    ref = (A2.float() @ B2.float().t())
    name = f"M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn} epi={epi} splits={splits}"
    ok = True
    if epi == 0:
        D0, _ = gemm(A, a_mn, B, b_mn, M, N, K, 0, alpha=0.5, bias=bias)
        ok = check(name, D0, 0.5 * ref + bias, 1e-2)
    elif epi == 1:
        D0, D1 = gemm(A, a_mn, B, b_mn, M, N, K, 1, alpha=0.125, bias=bias)
        pre = 0.125 * ref + bias
        ok = check(name + " pre", D0, pre, 1e-2)
        ok &= check(name + " act", D1, quick_gelu(D0.float()), 1e-2)
    elif epi == 2:
        aux = torch.randn(M, N, device=dev).bfloat16()
        D0, _ = gemm(A, a_mn, B, b_mn, M, N, K, 2, alpha=0.125, aux=aux)
        x = aux.float()
        s = torch.sigmoid(1.702 * x)
        ok = check(name, D0, 0.125 * ref * (s * (1 + 1.702 * x * (1 - s))), 1e-2)
    else:
        D0, _ = gemm(A, a_mn, B, b_mn, M, N, K, 3, alpha=1.0, bias=bias, splits=splits)
        ok = check(name, D0, ref + bias, 2e-5 * K ** 0.5 + 1e-5)
        D0b, _ = gemm(A, a_mn, B, b_mn, M, N, K, 3, alpha=1.0, bias=None, splits=splits, accumulate=1, D0=D0.clone())
        ok &= check(name + " accumulate", D0b, 2 * ref + bias, 2e-5 * K ** 0.5 + 1e-5)
    torch.cuda.synchronize()
    return ok


def bench(M, N, K, a_mn, b_mn, epi, splits=1, iters=10, name=""):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    aux = torch.randn(M, N, device=dev).bfloat16() if epi == 2 else None
    bias = torch.randn(N, device=dev)
    D0 = torch.empty((M, N), device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16)
    D1 = torch.empty((M, N), device=dev, dtype=torch.bfloat16) if epi == 1 else None
    for _ in range(3):
        gemm(A, a_mn, B, b_mn, M, N, K, epi, bias=bias, aux=aux, splits=splits, D0=D0, D1=D1, accumulate=int(epi == 3))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gemm(A, a_mn, B, b_mn, M, N, K, epi, bias=bias, aux=aux, splits=splits, D0=D0, D1=D1, accumulate=int(epi == 3))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS reference point (library baseline, not the product)
    A2 = A.t() if a_mn else A
    B2 = B.t() if b_mn else B
    Dc = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(A2, B2.t(), out=Dc)
    e0.record()
    for _ in range(iters):
        torch.matmul(A2, B2.t(), out=Dc)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    log(f"BENCH {name} M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn} epi={epi} splits={splits}: {ms:.3f} ms {tf:.1f} TFLOP/s | cuBLAS {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.1f} TFLOP/s")


def main():
    log("device:", torch.cuda.get_device_name(0))
    all_ok = True
    cases = [
        (128, 256, 64, 0, 0, 3, 1), (256, 512, 256, 0, 0, 3, 1), (1000, 768, 200, 0, 0, 3, 1),
        (384, 512, 512, 0, 0, 0, 1), (1000, 768, 328, 0, 0, 0, 1), (512, 1024, 256, 0, 0, 1, 1),
        (256, 256, 128, 0, 1, 3, 1), (1000, 768, 264, 0, 1, 0, 1), (640, 512, 512, 0, 1, 2, 1),
        (256, 256, 128, 1, 1, 3, 1), (768, 768, 4096, 1, 1, 3, 4), (1000, 520, 1000, 1, 1, 3, 3),
        (256, 256, 128, 1, 0, 3, 1),
        (4096, 2304, 768, 0, 0, 0, 1), (5000, 3072, 768, 0, 1, 2, 1), (5000, 3072, 768, 0, 0, 1, 1),
    ]
    for c in cases:
        try:
            all_ok &= run_case(*c)
        except Exception as e:  # noqa: BLE001
            log("EXC", c, repr(e))
            all_ok = False
            break
    log("ALL_OK" if all_ok else "SOME_FAILED")
    if all_ok or os.environ.get("BENCH_ANYWAY"):
        T = 1024 * 197
        bench(T, 2304, 768, 0, 0, 0, name="qkv_fwd")
        bench(T, 768, 768, 0, 0, 0, name="out_fwd")
        bench(T, 3072, 768, 0, 0, 1, name="fc1_fwd")
        bench(T, 768, 3072, 0, 0, 0, name="fc2_fwd")
        bench(T, 3072, 768, 0, 1, 2, name="fc2_dgrad")
        bench(T, 768, 3072, 0, 1, 0, name="fc1_dgrad")
        bench(3072, 768, T, 1, 1, 3, splits=8, name="fc1_wgrad")
        bench(768, 3072, T, 1, 1, 3, splits=8, name="fc2_wgrad")
        bench(2304, 768, T, 1, 1, 3, splits=8, name="qkv_wgrad")
        bench(8192, 8192, 8192, 0, 0, 0, name="square")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_probe.log", "w") as f:
        f.write("\n".join(out_lines) + "\n")


if __name__ == "__main__":
    main()
