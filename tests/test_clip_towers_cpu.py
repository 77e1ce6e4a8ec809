"""Host logic of the benchmarked hot path WITHOUT a GPU: the CLIP towers' fused forward / backward schedule
(engine.ViTTower / TextTower / TransformerStack behind autograd.TowerFunction) on the emulated kernel contracts
(tests/emu_ops.py) against autograd over the fp32 oracle (oracle/clip_oracle.py) — every parameter gradient.  The kernels
themselves and the full-size step are checked on the GPU (tests/test_gpu_parity.py); this test protects the schedule
(buffer routing, gradient slots, fused bias-gradient sums, gather-mode LayerNorms) on the CPU-only CI leg."""
import pytest
import torch

import emu_ops
from oracle import clip_oracle as O


@pytest.fixture()
def emu(monkeypatch):
    emu_ops.install(monkeypatch)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def test_clip_towers_schedule_against_oracle_with_emulated_kernels(emu):
    from multimodal_b200.models.clip.image_encoder import CLIPViTEncoder
    from multimodal_b200.models.clip.model import CLIP
    from multimodal_b200.models.clip.text_encoder import CLIPTextEncoder

    torch.manual_seed(0)
    m = CLIP(CLIPViTEncoder(64, 16, 64, 128, 2, 2),
             CLIPTextEncoder(embedding_dim=64, vocab_size=512, width=128, dim_feedforward=512, heads=2, layers=2)).train()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.03 * torch.randn(p.shape, generator=g))
    B = 5
    image = torch.randn(B, 3, 64, 64, generator=g)
    text = torch.randint(1, 500, (B, 77), generator=g)
    text[torch.arange(B), torch.randint(5, 77, (B,), generator=g)] = 511          # EOT = the largest id
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
    ra, rb = O.clip_forward(image, text, sd, 2, 2)
    wa, wb = torch.randn(ra.shape, generator=g), torch.randn(rb.shape, generator=g)
    ((wa * ra).sum() + (wb * rb).sum()).backward()
    out = m(image, text)
    assert _rel(out.embeddings_a, ra) < 2e-2 and _rel(out.embeddings_b, rb) < 2e-2
    ((wa * out.embeddings_a).sum() + (wb * out.embeddings_b).sum()).backward()
    rows = []
    for k, p in m.named_parameters():
        ref = sd[k].grad
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        if ref.norm().item() == 0.0:
            assert p.grad.abs().max().item() < 1e-6, k
            continue
        if k.endswith("in_proj_bias"):     # the key third is exactly zero in exact arithmetic: compare q / v thirds
            d = ref.numel() // 3
            rows.append((k + "[q]", _rel(p.grad[:d], ref[:d])))
            rows.append((k + "[v]", _rel(p.grad[2 * d:], ref[2 * d:])))
            continue
        rows.append((k, _rel(p.grad, ref)))
    worst = sorted(rows, key=lambda r: -r[1])[:5]
    print(worst)
    assert len(rows) > 60
    for k, e in rows:
        assert e < 5e-2, (k, e)
