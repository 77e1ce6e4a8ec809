"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference, which exists only
in the build container) on small deterministic inputs.  Committed together with its outputs; the GPU box only ever
reads the .pt files.

    PYTHONPATH=oracle/iopath_shim:/root/reference python tests/golden/make_golden.py

Cases (kept tiny so the fixtures stay small; weights are regenerated from seeds, only inputs/outputs are stored):
  clip_small : CLIP(CLIPViTEncoder(64, 16, 64, 128, 2, 2), CLIPTextEncoder(64, vocab=512, width=128, ff=512, heads=2, layers=2))
               B=4, image 64x64, reference init under torch.manual_seed(0); fp32 forward, loss, and all gradients
               (stored as per-tensor float64 sums + a few raw tensors).
  clip_b16_b2: clip_vit_b16() under torch.manual_seed(0), B=2 synthetic batch; embeddings + loss only.
"""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "iopath_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from torchmultimodal.models.clip.image_encoder import CLIPViTEncoder  # noqa: E402
from torchmultimodal.models.clip.model import CLIP, clip_vit_b16  # noqa: E402
from torchmultimodal.models.clip.text_encoder import CLIPTextEncoder  # noqa: E402
from torchmultimodal.modules.losses.contrastive_loss_with_temperature import (  # noqa: E402
    ContrastiveLossWithTemperature,
    contrastive_loss_with_temperature,
)

from oracle import clip_oracle as O  # noqa: E402


def small_model():
    torch.manual_seed(0)
    enc_a = CLIPViTEncoder(embedding_dim=64, patch_size=16, image_size=64, width=128, heads=2, layers=2)
    enc_b = CLIPTextEncoder(embedding_dim=64, vocab_size=512, width=128, dim_feedforward=512, heads=2, layers=2)
    return CLIP(enc_a, enc_b)


def main():
    torch.set_num_threads(8)
    out = {}
    # ---------------- clip_small: forward + loss + grads ----------------
    m = small_model().train()
    with torch.no_grad():  # break the "all layers identical" symmetry of nn.TransformerEncoder's deepcopy init
        g = torch.Generator().manual_seed(7)
        for p in m.parameters():
            p.add_(0.02 * torch.randn(p.shape, generator=g))
    img, txt = O.synthetic_batch(4, image_size=64, vocab=512)
    loss_mod = ContrastiveLossWithTemperature()
    o = m(img, txt)
    res = contrastive_loss_with_temperature(o.embeddings_a, o.embeddings_b, loss_mod.logit_scale)
    res.loss.backward()
    out["clip_small"] = {
        "image": img, "text": txt,
        "state_dict": {k: v.clone() for k, v in m.state_dict().items()},
        "emb_a": o.embeddings_a.detach().clone(), "emb_b": o.embeddings_b.detach().clone(),
        "loss": res.loss.detach().clone(), "logits_a": res.logits_a.detach().clone(),
        "logits_b": res.logits_b.detach().clone(),
        # full gradients for small tensors; (abs-sum, abs-max, 4096 evenly strided samples) for the large ones
        "grads": {k: (p.grad.detach().clone() if p.numel() <= 20000 else
                      {"abssum": p.grad.double().abs().sum(), "absmax": p.grad.abs().max(),
                       "stride": max(1, p.numel() // 4096),
                       "sample": p.grad.reshape(-1)[::max(1, p.numel() // 4096)].clone(), "shape": tuple(p.shape)})
                  for k, p in m.named_parameters()},
        "logit_scale_grad": loss_mod.logit_scale.grad.detach().clone(),
    }
    # ---------------- clip_b16_b2: full-size forward ----------------
    torch.manual_seed(0)
    big = clip_vit_b16().eval()
    img, txt = O.synthetic_batch(2)
    with torch.no_grad():
        big.train()  # python MHA path (SURVEY.md §7 'three numerics paths'); dropout is 0
        ob = big(img, txt)
        lb = contrastive_loss_with_temperature(ob.embeddings_a, ob.embeddings_b, torch.tensor(math.log(1 / 0.07)))
    out["clip_b16_b2"] = {"emb_a": ob.embeddings_a.clone(), "emb_b": ob.embeddings_b.clone(), "loss": lb.loss.clone(),
                          "logits_a": lb.logits_a.clone()}
    torch.save(out, os.path.join(HERE, "clip_golden.pt"))
    print("wrote", os.path.join(HERE, "clip_golden.pt"), os.path.getsize(os.path.join(HERE, "clip_golden.pt")) / 1e6, "MB")


if __name__ == "__main__":
    main()
