"""Timing probe: one FLAVA pre-training step (BASELINE.json config 3 as a TRAINING step) — FLAVAForPreTraining forward
(image / text encoders twice, multimodal encoder, ITM + MMM heads + global contrastive loss) + backward on the library's
kernels, full-size model (12 / 12 / 6 layers, d = 768, 224x224 images, 77 tokens, vocabularies 30 522 / 8 192).
Run under gpurun; writes gpurun_out/flava_train_probe.log."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import flava_pretraining_cases as PC  # noqa: E402
from multimodal_b200 import ops  # noqa: E402
from multimodal_b200.models.flava import flava_model, FLAVAForPreTraining  # noqa: E402
from multimodal_b200.modules.losses.flava import FLAVAPretrainingLoss  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B = int(os.environ.get("FLAVA_BS", "128"))
    torch.manual_seed(0)
    m = FLAVAForPreTraining(model=flava_model(), image_codebook=PC.StubCodebook(16, 8192),
                            loss=FLAVAPretrainingLoss()).to(dev).train()
    g = torch.Generator().manual_seed(1)
    image = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    text = torch.randint(1, 30000, (B, 77), generator=g)
    text[:, 60:] = 0
    text_masked = text.clone()
    mlm = torch.full((B, 77), -1, dtype=torch.long)
    pick = (torch.rand(B, 77, generator=g) < 0.15) & (text != 0)
    mlm[pick] = text[pick]
    text_masked[pick] = 103
    pm = torch.rand(B, 14, 14, generator=g) < 0.4
    itm = (torch.rand(B, generator=g) < 0.9).long()
    inp = dict(image=image, text=text.to(dev), image_for_codebook=image, image_patches_mask=pm.to(dev),
               text_masked=text_masked.to(dev), itm_labels=itm.to(dev), mlm_labels=mlm.to(dev))
    opt = torch.optim.SGD(m.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        out = m(**inp)
        total = sum(v for v in out.losses.values() if v is not None)
        total.backward()
        opt.step()
        return total

    lines = []
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
    lines.append(f"FLAVAForPreTraining train step bs={B}: {ms:.1f} ms/step = {B / ms * 1e3:.0f} samples/s; loss {t.item():.4f}; "
                 f"GEMM kernels {gemm_ms:.1f} ms, {gemm_fl / 1e12:.1f} TFLOP -> {gemm_fl / gemm_ms / 1e9:.0f} TFLOP/s "
                 f"({len(gt)} launches); peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    print("\n".join(lines), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/flava_train_probe.log", "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
