"""Timing probe of the GPU CLIP image transform (image half): N decoded 375x500 RGB images already in HBM ->
[N, 3, 224, 224] fp32, eval (Resize + CenterCrop) and train (RandomResizedCrop) modes; algorithmic bytes = source pixels
read once + output written once.  CPU reference point: the same pipeline through Pillow / torchvision on one host core.
Run under gpurun; writes gpurun_out/clip_transform_probe.log."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_b200.transforms.clip_transform import CLIPImageTransform  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    N, H, W, S = int(os.environ.get("N_IMAGES", "256")), 375, 500, 224
    rng = np.random.default_rng(0)
    host = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(8)]
    imgs = [torch.from_numpy(host[i % 8]).to(dev) for i in range(N)]
    lines = []
    for mode in (False, True):
        t = CLIPImageTransform(image_size=S, is_train=mode)
        torch.manual_seed(0)
        for _ in range(3):
            out = t(imgs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = t(imgs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nbytes = N * (H * W * 3 + 3 * S * S * 4)
        lines.append(f"GPU CLIPImageTransform is_train={mode}: {N} x {H}x{W} -> {S}x{S}: {ms:.3f} ms per batch (host geometry + "
                     f"2 kernels) = {N / ms * 1e3:.0f} images/s, {nbytes / ms / 1e6:.1f} GB/s of algorithmic bytes")
    try:
        from PIL import Image
        from torchvision import transforms as T
        from torchvision.transforms import InterpolationMode
        torch.set_num_threads(1)
        tv = T.Compose([T.Resize(S, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(S), T.ToTensor(),
                        T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])
        pil = [Image.fromarray(h) for h in host]
        t0 = time.perf_counter()
        for i in range(64):
            tv(pil[i % 8])
        dt = time.perf_counter() - t0
        lines.append(f"CPU reference pipeline (Pillow + torchvision, 1 core, same images, eval): {64 / dt:.0f} images/s")
    except Exception as e:  # noqa: BLE001
        lines.append(f"CPU reference pipeline not timed: {e}")
    print("\n".join(lines), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/clip_transform_probe.log", "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
