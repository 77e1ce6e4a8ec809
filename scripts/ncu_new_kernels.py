"""One launch of each kernel added in the re-entry session at a realistic shape, for `ncu --set full` (and, run plainly,
CUDA-event timings of the same launches): the CLIP image transform (256 decoded 375x500 images -> 224x224) and the
general attention backward at CoCa ViT-L/14 shapes (captioning pooler: 256 shared queries x 257 keys, head_dim 96;
multimodal decoder cross-attention: 76 x 256, head_dim 64), B = 64.  Writes gpurun_out/new_kernels_timing.log when run
without ncu (NCU=0)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_b200 import ops  # noqa: E402
from multimodal_b200.transforms.clip_transform import CLIPImageTransform  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    under_ncu = os.environ.get("NCU", "0") == "1"
    lines = []
    # ---- image transform: kernels only (geometry / pointer tensors prepared once)
    N, H, W, S = 256, 375, 500, 224
    rng = np.random.default_rng(0)
    host = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(8)]
    imgs = [torch.from_numpy(host[i % 8]).to(dev) for i in range(N)]
    t = CLIPImageTransform(image_size=S, is_train=False)
    geom = torch.tensor([t._geometry(H, W, W * 3) for _ in range(N)], dtype=torch.int32, device=dev)
    ptrs = torch.tensor([im.data_ptr() for im in imgs], dtype=torch.int64, device=dev)
    out = torch.empty((N, 3, S, S), device=dev)
    run_t = lambda: ops.clip_image_transform(ptrs, geom, out, t.image_mean, t.image_std)  # noqa: E731
    # ---- general attention backward
    B = 64
    bf = torch.bfloat16

    def attn_case(Sq, Skv, Hh, hd, shared):
        d = Hh * hd
        q = (torch.randn(Sq if shared else B * Sq, d, device=dev) * 0.5).to(bf)
        kv = (torch.randn(B * Skv, 2 * d, device=dev) * 0.5).to(bf)
        do = (torch.randn(B * Sq, d, device=dev) * 0.5).to(bf)
        dkv = torch.empty_like(kv)
        dq = None if shared else torch.empty_like(q)
        dq32 = torch.zeros(Sq, d, device=dev) if shared else None
        kw = dict(B=B, Sq=Sq, Skv=Skv, H=Hh, head_dim=hd, bsq=0 if shared else Sq * d, bsk=Skv * 2 * d, bsv=Skv * 2 * d,
                  bso=Sq * d, scale=1.0 / math.sqrt(hd))
        return lambda: ops.attention_bwd_generic(q, kv[:, :d], kv[:, d:], do, dkv[:, :d], dkv[:, d:], dq=dq, dq_f32=dq32, **kw)

    run_pool = attn_case(256, 257, 8, 96, True)
    run_cross = attn_case(76, 256, 12, 64, False)
    if under_ncu:
        for fn in (run_t, run_pool, run_cross):
            fn()
        torch.cuda.synchronize()
        return
    ms = timed(run_t)
    nbytes = N * (H * W * 3 + 3 * S * S * 4)
    lines.append(f"clip image transform kernels (coeffs + resample), {N} x {H}x{W} -> {S}x{S}: {ms:.3f} ms = {N / ms * 1e3:.0f} "
                 f"images/s, {nbytes / ms / 1e6:.0f} GB/s of algorithmic bytes ({nbytes / 1e6:.0f} MB)")
    for name, fn, Sq, Skv, Hh, hd in (("pooler 256 shared queries x 257 keys, hd 96, H 8", run_pool, 256, 257, 8, 96),
                                     ("cross-attention 76 x 256, hd 64, H 12", run_cross, 76, 256, 12, 64)):
        ms = timed(fn)
        fl = 10.0 * Sq * Skv * hd * Hh * B
        lines.append(f"general attention backward (SIMT), B = {B}, {name}: {ms:.3f} ms ({fl / ms / 1e9:.1f} TFLOP/s of the "
                     f"algorithmic 10*Sq*Skv*hd flops)")
    print("\n".join(lines), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/new_kernels_timing.log", "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
