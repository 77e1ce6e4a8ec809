"""GPU: the layer modules the reference exposes (and tests) on their own — MultiHeadSelfAttention, TransformerEncoderLayer
(pre- / post-norm), TransformerEncoder, PatchEmbeddings, MLP, SiLU, Fp32LayerNorm — called STANDALONE on the CUDA path
(engine_layers.py: same kernels as the fused encoders) against a plain fp32 restatement of the reference forward:
  modules/layers/multi_head_attention.py:39-80, modules/layers/transformer.py:76-154,216-259,
  modules/layers/patch_embedding.py:104-154, modules/layers/mlp.py:62-66, modules/layers/activation.py:24-25,
  modules/layers/normalizations.py:17-25.
The reference's own known-answer tests for these layers use 2..8-wide toy dimensions (tests/modules/layers/
test_multi_head_attention.py:32-48, test_transformer.py:45-189, test_patch_embedding.py:60-134) that no tensor-core
tile can hold (head_dim 64, feature sizes multiples of 8); the same forward definitions are checked here at real sizes.
Tolerance: bf16 GEMM operands, fp32 accumulation / LayerNorm / softmax statistics -> 2e-2 of the tensor's absmax.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-20)).item()


def _mhsa_ref(m, x, mask=None, causal=False):
    B, S, d = x.shape
    H = m.num_heads
    q, k, v = F.linear(x, m.input_proj.weight, m.input_proj.bias).chunk(3, dim=-1)
    q, k, v = (t.view(B, S, H, d // H).transpose(1, 2) for t in (q, k, v))
    s = q @ k.transpose(-1, -2) / math.sqrt(d // H)
    if causal:
        s = s + torch.full((S, S), float("-inf"), device=x.device).triu(1)
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, d)
    return F.linear(o, m.output_proj.weight, m.output_proj.bias)


@pytest.mark.parametrize("B,S,d,H,causal,masked", [(2, 197, 768, 12, False, False), (3, 77, 512, 8, True, False),
                                                 (2, 50, 256, 4, False, True), (2, 40, 192, 2, False, False)])
def test_multi_head_self_attention_standalone(B, S, d, H, causal, masked):
    from multimodal_b200._lib import MMBError
    from multimodal_b200.modules.layers.multi_head_attention import MultiHeadSelfAttention

    torch.manual_seed(0)
    m = MultiHeadSelfAttention(d, H).to(dev)
    x = torch.randn(B, S, d, device=dev)
    mask = None
    if masked:
        mask = torch.rand(B, 1, S, S, device=dev) < 0.7
        mask[..., 0] = True
    with pytest.raises(MMBError):       # forward values only: asking for a graph must fail loudly
        m(x, mask, causal)
    with torch.no_grad():
        got = m(x, mask, causal)
        ref = _mhsa_ref(m, x, mask, causal)
    assert got.shape == ref.shape and _rel(got, ref) < 2e-2, _rel(got, ref)


def _layer_ref(l, x, mask=None):
    def attn(h):
        return _mhsa_ref(l.attention, h, mask[:, None] if mask is not None else None)

    def ff(h):
        return l.feedforward.model(h) if False else F.linear(F.gelu(F.linear(h, l.feedforward.model[0].weight,
                                                                            l.feedforward.model[0].bias)),
                                                               l.feedforward.model[-1].weight, l.feedforward.model[-1].bias)

    def ln(mod, h):
        return F.layer_norm(h, h.shape[-1:], mod.weight, mod.bias, mod.eps)

    if l.norm_first:                                   # transformer.py:95-111
        a = attn(ln(l.attention_layernorm, x)) + x
        return a + ff(ln(l.feedforward_layernorm, a))
    a = ln(l.attention_layernorm, attn(x) + x)         # :113-128
    return ln(l.feedforward_layernorm, a + ff(a))


@pytest.mark.parametrize("norm_first", [True, False])
@pytest.mark.parametrize("masked", [False, True])
def test_transformer_encoder_layer_and_stack_standalone(norm_first, masked):
    from multimodal_b200.modules.layers.transformer import TransformerEncoder, TransformerEncoderLayer

    torch.manual_seed(1)
    B, S, d, H, ff = 3, 50, 256, 4, 1024
    layer = TransformerEncoderLayer(d, H, ff, activation=torch.nn.GELU, layer_norm_eps=1e-5, norm_first=norm_first).to(dev)
    with torch.no_grad():
        for p in layer.parameters():                   # away from the default init so that biases / affine terms matter
            p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(B, S, d, device=dev)
    mask = (torch.rand(B, S, S, device=dev) < 0.8) if masked else None
    if mask is not None:
        mask[..., 0] = True
    with torch.no_grad():
        got = layer(x, mask)
        ref = _layer_ref(layer, x, mask)
    assert _rel(got, ref) < 2e-2, (norm_first, masked, _rel(got, ref))
    if not masked:
        enc = TransformerEncoder(2, d, H, ff, activation=torch.nn.GELU, layer_norm_eps=1e-5, norm_first=norm_first,
                                 final_layer_norm_eps=1e-5).to(dev)
        with torch.no_grad():
            out = enc(x, return_hidden_states=True)
            h = x
            for l in enc.layer:
                h = _layer_ref(l, h)
            h = F.layer_norm(h, (d,), enc.final_layer_norm.weight, enc.final_layer_norm.bias, 1e-5)
        assert len(out.hidden_states) == 3 and _rel(out.last_hidden_state, h) < 3e-2


@pytest.mark.parametrize("cls,use_mask", [(True, False), (False, False), (True, True)])
def test_patch_embeddings_standalone(cls, use_mask):
    from multimodal_b200.modules.layers.patch_embedding import PatchEmbeddings

    torch.manual_seed(2)
    pe = PatchEmbeddings(image_size=64, patch_size=16, hidden_size=128, use_image_masking=use_mask, include_cls_embed=cls).to(dev)
    with torch.no_grad():
        pe.position_embeddings.normal_(0, 0.02)
        pe.conv_projection.bias.normal_(0, 0.02)
        if cls:
            pe.cls_token.normal_(0, 0.02)
        if use_mask:
            pe.mask_token.normal_(0, 0.02)
    img = torch.randn(5, 3, 64, 64, device=dev)
    pm = (torch.rand(5, 16, device=dev) < 0.4) if use_mask else None
    with torch.no_grad():
        got = pe(img, pm).embeddings
        e = pe.conv_projection(img).flatten(2).transpose(1, 2)            # patch_embedding.py:118-121
        if pm is not None:
            e = torch.where(pm[..., None], pe.mask_token.expand_as(e), e)  # :127-133
        if cls:
            e = torch.cat([pe.cls_token.expand(5, -1, -1), e], dim=1)      # :142-146
        ref = e + pe.position_embeddings                                    # :149
    assert got.shape == ref.shape and _rel(got, ref) < 1e-2, _rel(got, ref)


def test_activation_layernorm_mlp_standalone():
    from multimodal_b200.modules.layers.activation import SiLU
    from multimodal_b200.modules.layers.mlp import MLP
    from multimodal_b200.modules.layers.normalizations import Fp32LayerNorm

    torch.manual_seed(3)
    x = torch.randn(7, 33, 256, device=dev)
    torch.testing.assert_close(SiLU()(x), torch.sigmoid(1.702 * x) * x, rtol=1e-5, atol=1e-6)
    # the reference's known answer (tests/modules/layers/test_activation.py:12-16): SiLU(1) = 0.8458
    assert abs(SiLU()(torch.ones(1, device=dev)).item() - 0.8458) < 1e-4
    ln = Fp32LayerNorm(256).to(dev)   # the LayerNorm kernels take widths that are multiples of 128
    with torch.no_grad():
        ln.weight.normal_(1, 0.1); ln.bias.normal_(0, 0.1)
        got = ln(x.bfloat16())
        assert got.dtype == torch.bfloat16                                  # type_as(x) (normalizations.py:25)
        ref = F.layer_norm(x.bfloat16().float(), (256,), ln.weight, ln.bias, ln.eps)
    assert _rel(got, ref) < 1e-2
    mlp = MLP(256, 128, 512, dropout=0.0, activation=torch.nn.GELU).to(dev)
    y = torch.randn(40, 256, device=dev)
    with torch.no_grad():
        assert _rel(mlp(y), mlp.model(y)) < 2e-2
