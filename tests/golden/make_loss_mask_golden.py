"""Generates tests/golden/loss_mask_golden.pt: the UNMODIFIED reference `contrastive_loss_with_temperature` with the
row `mask` argument (contrastive_loss_with_temperature.py:97-100), single process, forward + autograd gradients.

    python tests/golden/make_loss_mask_golden.py
"""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "iopath_shim"))
sys.path.insert(0, "/root/reference")

from torchmultimodal.modules.losses.contrastive_loss_with_temperature import (  # noqa: E402
    contrastive_loss_with_temperature,
)


def case(B, E, seed, smoothing, normalize):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, E, generator=g)
    b = torch.randn(B, E, generator=g)
    if normalize:
        a, b = torch.nn.functional.normalize(a, dim=1), torch.nn.functional.normalize(b, dim=1)
    mask = torch.rand(B, generator=g) < 0.7
    mask[0] = True
    a.requires_grad_(True), b.requires_grad_(True)
    s = torch.nn.Parameter(torch.tensor(math.log(1 / 0.07)))
    out = contrastive_loss_with_temperature(a, b, s, mask=mask,
                                            cross_entropy_kwargs={"label_smoothing": smoothing} if smoothing else None)
    out.loss.backward()
    return dict(a=a.detach(), b=b.detach(), mask=mask, smoothing=smoothing, loss=out.loss.detach(),
                loss_a=out.loss_a.detach(), loss_b=out.loss_b.detach(), logits_a=out.logits_a.detach(),
                logits_b=out.logits_b.detach(), dA=a.grad, dB=b.grad, dS=s.grad)


def main():
    out = {"kat_3x5": case(3, 5, 1234, 0.0, False), "b128_e64": case(128, 64, 0, 0.1, True)}
    path = os.path.join(HERE, "loss_mask_golden.pt")
    torch.save(out, path)
    print({k: (float(v["loss"]), int(v["mask"].sum())) for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
