"""Timing probe: one CoCa pre-training step (BASELINE.json config 5 as a TRAINING step) — `CoCaForPretraining` forward
(ViT-L/14 vision encoder, cascaded attention poolers, text decoder, multimodal decoder, contrastive + captioning losses),
backward and an SGD update on the library's kernels, `coca_vit_l_14()` (584 M parameters), 224x224 images, 77 tokens.
Run under gpurun; writes gpurun_out/coca_train_probe.log."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_b200 import ops  # noqa: E402
from multimodal_b200.models.coca.coca_model import CoCaForPretraining, coca_vit_l_14  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B = int(os.environ.get("COCA_BS", "64"))
    torch.manual_seed(0)
    m = CoCaForPretraining(coca_vit_l_14()).to(dev).train()
    g = torch.Generator().manual_seed(1)
    images = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    texts = torch.randint(1, 49000, (B, 77), generator=g)
    texts[:, 50:] = 0
    texts = texts.to(dev)
    opt = torch.optim.SGD(m.parameters(), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        out = m(images, texts)
        total = out["contrastive"] + out["captioning"]
        total.backward()
        opt.step()
        return total

    for _ in range(2):
        t = step()
    torch.cuda.synchronize()
    n = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        t = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ops.GEMM_TIMING = []
    step()
    torch.cuda.synchronize()
    gt = ops.GEMM_TIMING
    ops.GEMM_TIMING = None
    gemm_ms = sum(ev[0].elapsed_time(ev[1]) for _, _, ev in gt)
    gemm_fl = sum(f for f, _, _ in gt)
    line = (f"CoCaForPretraining (coca_vit_l_14) train step bs={B}: {ms:.1f} ms/step = {B / ms * 1e3:.0f} samples/s; loss "
            f"{t.item():.4f}; GEMM kernels {gemm_ms:.1f} ms, {gemm_fl / 1e12:.1f} TFLOP -> {gemm_fl / gemm_ms / 1e9:.0f} TFLOP/s "
            f"({len(gt)} launches); peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    print(line, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/coca_train_probe.log", "w").write(line + "\n")


if __name__ == "__main__":
    main()
