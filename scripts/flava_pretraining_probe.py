"""Timing probe of the FLAVA pre-training losses at the real head shapes (hidden 768, text vocabulary 30 522, image
vocabulary 8 192; BASELINE config 3 batch 256: 196 patches + CLS, 128 text tokens, 15 % / 40 % of the tokens masked).
CUDA events, 10 iterations after warm-up; prints ms per call and the GEMM TFLOP/s of the vocabulary projections."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_b200.modules.losses.flava import FLAVAPretrainingLoss  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S_IMG, S_TXT, d = 256, 197, 128, 768
loss = FLAVAPretrainingLoss(hidden_size=d).to(dev).eval()
g = torch.Generator(device="cpu").manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g).to(dev)  # noqa: E731
mlm = torch.full((B, S_TXT), -1, dtype=torch.long)
mim = torch.full((B, S_IMG - 1), -1, dtype=torch.long)
pt, pi = torch.rand(B, S_TXT, generator=g) < 0.15, torch.rand(B, S_IMG - 1, generator=g) < 0.4
mlm[pt] = torch.randint(0, 30522, (int(pt.sum()),), generator=g)
mim[pi] = torch.randint(0, 8192, (int(pi.sum()),), generator=g)
kw = dict(image_sequence=r(B, S_IMG, d), text_sequence=r(B, S_TXT, d), image_masked_sequence=r(B, S_IMG, d),
          text_masked_sequence=r(B, S_TXT, d), multimodal_masked_sequence=r(B, 1 + S_IMG + S_TXT, d),
          itm_labels=(torch.rand(B, generator=g) < 0.7).long().to(dev), mlm_labels=mlm.to(dev), mim_labels=mim.to(dev),
          projected_image_embeddings=r(B, d), projected_text_embeddings=r(B, d))
uni = {k: v for k, v in kw.items() if k not in ("multimodal_masked_sequence", "itm_labels")}


def t(fn, iters=10):
    with torch.no_grad():
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


n_t, n_i = int(pt.sum()), int(pi.sum())
ms_u = t(lambda: loss(**uni))
ms_m = t(lambda: loss(**kw))
fl = 2.0 * d * (n_t * (30522 + d) + n_i * (8192 + d))
print(f"FLAVAPretrainingLoss bs={B}: unimodal branch (MLM {n_t} rows x 30522 + MIM {n_i} rows x 8192 + contrastive) "
      f"{ms_u:.2f} ms ({fl / ms_u / 1e9:.0f} TFLOP/s over the head GEMMs); multimodal branch (ITM + MMM text / image + "
      f"contrastive over the positive pairs) {ms_m:.2f} ms")
