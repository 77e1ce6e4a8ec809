"""GPU probe: CoCa ViT-L/14 forward (+ both pretraining losses) timing at a small batch.  Writes gpurun_out/coca_probe.log."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_b200.models.coca import coca_vit_l_14, CoCaForPretraining  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("COCA_BS", "64"))
t0 = time.time()
torch.manual_seed(0)
m = CoCaForPretraining(coca_vit_l_14()).to(dev).eval()
print(f"model built in {time.time() - t0:.1f} s, {sum(p.numel() for p in m.parameters()) / 1e6:.1f} M params", flush=True)
images = torch.randn(B, 3, 224, 224, device=dev)
texts = torch.randint(1, 49408, (B, 77), device=dev)


def timed(fn, n=3, warm=1):
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


t_model = timed(lambda: m.model(images, texts))
t_vit = timed(lambda: m.model.vision_encoder(images))
torch.set_grad_enabled(False)   # forward-only runtime (CoCaForPretraining refuses trainable losses)
t_all = timed(lambda: m(images, texts))
gf = 205.9  # GF per sample forward
line = (f"CoCa ViT-L/14 forward bs={B}: vision encoder {t_vit:.2f} ms | CoCaModel.forward {t_model:.2f} ms = {B / t_model * 1e3:.0f} "
        f"samples/s ({gf * B / t_model:.0f} TFLOP/s model-level) | + contrastive & captioning losses {t_all:.2f} ms")
print(line, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/coca_probe.log", "w").write(line + "\n")
