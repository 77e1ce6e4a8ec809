"""GPU parity of the CoCa backward building blocks (SURVEY.md §8 a14 / f3 as training steps) against torch autograd."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,Sq,Skv,H,hd,kind", [
    (3, 13, 37, 2, 64, "cross"),          # multimodal decoder cross-attention
    (2, 76, 256, 12, 64, "cross"),        # ... at CoCa's real sequence lengths
    (3, 5, 20, 2, 96, "shared_q"),        # attention pooler: learned queries shared by the batch, head_dim 96
    (2, 256, 257, 8, 96, "shared_q"),     # ... captioning pooler of ViT-L/14
    (4, 1, 40, 2, 128, "shared_q"),       # contrastive pooler (one query), head_dim 128
    (3, 21, 21, 2, 64, "mask"),           # text decoder: [B, S, S] boolean mask on a packed QKV buffer
    (2, 33, 33, 3, 64, "causal"),
])
def test_attention_bwd_generic(dev, B, Sq, Skv, H, hd, kind):
    from multimodal_b200 import ops

    torch.manual_seed(5)
    d = H * hd
    scale = 1.0 / math.sqrt(hd)
    bf = torch.bfloat16
    mask = None
    if kind in ("mask", "causal"):   # self-attention on a packed [B*S, 3d] buffer (column slices as operands)
        qkv = (torch.randn(B * Sq, 3 * d, device=dev) * 0.7).to(bf)
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
        bsq = bsk = bsv = Sq * 3 * d
        dqkv = torch.full_like(qkv, float("nan"))
        dq, dk, dv = dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:]
        if kind == "mask":
            mask = (torch.rand(B, Sq, Skv, device=dev) < 0.6)
            mask[:, :, 0] = True
            mask[1, 3, :] = False    # a fully masked query row: zeros, no NaN
    else:
        nq_rows = Sq if kind == "shared_q" else B * Sq
        q = (torch.randn(nq_rows, d, device=dev) * 0.7).to(bf)
        kv = (torch.randn(B * Skv, 2 * d, device=dev) * 0.7).to(bf)
        k, v = kv[:, :d], kv[:, d:]
        bsq = 0 if kind == "shared_q" else Sq * d
        bsk = bsv = Skv * 2 * d
        dkv = torch.full_like(kv, float("nan"))
        dk, dv = dkv[:, :d], dkv[:, d:]
        dq = None if kind == "shared_q" else torch.full_like(q, float("nan"))
    dout = (torch.randn(B * Sq, d, device=dev) * 0.5).to(bf)
    out = torch.empty(B * Sq, d, device=dev, dtype=bf)
    mu8 = mask.to(torch.uint8).contiguous() if mask is not None else None
    kw = dict(B=B, Sq=Sq, Skv=Skv, H=H, head_dim=hd, bsq=bsq, bsk=bsk, bsv=bsv, bso=Sq * d, scale=scale, mask=mu8,
              mask_bs=Sq * Skv if mask is not None else 0, mask_qs=Skv if mask is not None else 0, causal=kind == "causal")
    ops.attention_fwd_generic(q, k, v, out, **kw)
    dq32 = torch.zeros(Sq, d, device=dev) if kind == "shared_q" else None
    ops.attention_bwd_generic(q, k, v, dout, dk, dv, dq=dq, dq_f32=dq32, **kw)
    # reference
    qf = q.float().clone().requires_grad_(True)
    kf = k.float().clone().requires_grad_(True)
    vf = v.float().clone().requires_grad_(True)
    qh = (qf.view(1, Sq, H, hd).expand(B, Sq, H, hd) if kind == "shared_q" else qf.view(B, Sq, H, hd)).transpose(1, 2)
    kh, vh = kf.view(B, Skv, H, hd).transpose(1, 2), vf.view(B, Skv, H, hd).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    if kind == "causal":
        s = s + torch.full((Sq, Skv), float("-inf"), device=dev).triu(1)
    if mask is not None:
        s = s.masked_fill(~mask[:, None], float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)          # fully masked row -> zeros (the kernels' convention)
    ref = (p @ vh).transpose(1, 2).reshape(B * Sq, d)
    assert _rel(out, ref) < 1e-2
    ref.backward(dout.float())
    assert torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()
    assert _rel(dk, kf.grad) < 1e-2 and _rel(dv, vf.grad) < 1e-2
    if kind == "shared_q":
        assert _rel(dq32, qf.grad) < 1e-2
    else:
        assert torch.isfinite(dq.float()).all()
        assert _rel(dq, qf.grad) < 1e-2
