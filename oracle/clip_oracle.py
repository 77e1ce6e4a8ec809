"""ORACLE (test infrastructure — never imported by the product path).

A plain fp32 PyTorch restatement, op by op, of the reference's dual-encoder forward + contrastive loss.  It takes a
reference-format ``state_dict`` (same key names / shapes as ``torchmultimodal.models.clip.model.CLIP``) and raw
inputs, and uses only elementary tensor ops (matmul, softmax, mean/var, gather) so that every step can be read
against the reference line it restates.  Gradients come from autograd over these elementary ops.

Pinned (tests/test_oracle_cpu.py) against:
  * golden vectors produced by importing the UNMODIFIED reference from /root/reference in the build container
    (tests/golden/make_golden.py -> tests/golden/*.pt), and
  * the reference's own known-answer tests (values quoted in the test file with their file:line).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """F.layer_norm over the last dim (biased variance, eps inside sqrt).
    Restates: torch/nn/functional.py layer_norm as called by torch/nn/modules/transformer.py:946,951 (norm1/norm2)
    and torchmultimodal/modules/layers/normalizations.py:17-25 (Fp32LayerNorm: same math in fp32)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def quick_gelu(x: Tensor) -> Tensor:
    """torchmultimodal/modules/layers/activation.py:24-25 ('SiLU' == QuickGELU): sigmoid(1.702 x) * x."""
    return torch.sigmoid(1.702 * x) * x


def mha(x: Tensor, sd: Dict[str, Tensor], pfx: str, heads: int, causal: bool) -> Tensor:
    """Self-attention block. x: [B,S,d].
    Restates torch/nn/functional.py:6244 multi_head_attention_forward: packed in-projection (:6478), per-head
    softmax(q k^T / sqrt(dh)) v (:6682, causal mask when is_causal), out-projection (:6690)."""
    B, S, d = x.shape
    dh = d // heads
    qkv = x @ sd[pfx + "self_attn.in_proj_weight"].t() + sd[pfx + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(B, S, heads, dh).transpose(1, 2)
    k = k.view(B, S, heads, dh).transpose(1, 2)
    v = v.view(B, S, heads, dh).transpose(1, 2)
    att = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if causal:
        mask = torch.full((S, S), float("-inf"), device=x.device, dtype=x.dtype).triu(1)  # text_encoder.py:74-77
        att = att + mask
    att = torch.softmax(att, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, d)
    return o @ sd[pfx + "self_attn.out_proj.weight"].t() + sd[pfx + "self_attn.out_proj.bias"]


def encoder_layer(x: Tensor, sd: Dict[str, Tensor], pfx: str, heads: int, causal: bool) -> Tensor:
    """Pre-norm layer: torch/nn/modules/transformer.py:946-951 (_sa_block :961-978, _ff_block :980-982)."""
    h = layer_norm(x, sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"])
    x = x + mha(h, sd, pfx, heads, causal)
    h = layer_norm(x, sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"])
    h = quick_gelu(h @ sd[pfx + "linear1.weight"].t() + sd[pfx + "linear1.bias"])
    return x + (h @ sd[pfx + "linear2.weight"].t() + sd[pfx + "linear2.bias"])


def _num_layers(sd: Dict[str, Tensor], pfx: str) -> int:
    n = 0
    while f"{pfx}encoder.layers.{n}.norm1.weight" in sd:
        n += 1
    return n


def vit_encoder(image: Tensor, sd: Dict[str, Tensor], pfx: str, heads: int) -> Tensor:
    """models/clip/image_encoder.py:82-113."""
    w = sd[pfx + "conv.weight"]  # [width, 3, ps, ps]
    width, _, ps, _ = w.shape
    B, C, H, W = image.shape
    if C != 3:
        raise ValueError(f"Expected 3 channels found {C}")
    gh, gw = H // ps, W // ps
    # conv with stride == kernel == ps and no bias (:50-56,:91) == per-patch dot product, K order (c, kh, kw)
    patches = image.view(B, C, gh, ps, gw, ps).permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, C * ps * ps)
    x = patches @ w.reshape(width, -1).t()                                    # flatten(2).permute(0,2,1) :94-97
    cls = sd[pfx + "cls_token_embedding"].expand(B, 1, width)
    x = torch.cat([cls, x], dim=1) + sd[pfx + "positional_embedding"]         # :98-105
    x = layer_norm(x, sd[pfx + "ln_pre.weight"], sd[pfx + "ln_pre.bias"])    # :106
    for l in range(_num_layers(sd, pfx)):                                     # :108
        x = encoder_layer(x, sd, f"{pfx}encoder.layers.{l}.", heads, causal=False)
    x = layer_norm(x[:, 0, :], sd[pfx + "ln_post.weight"], sd[pfx + "ln_post.bias"])  # :111
    return x @ sd[pfx + "projection"]                                         # :112


def text_encoder(text: Tensor, sd: Dict[str, Tensor], pfx: str, heads: int, return_hidden_state: bool = False) -> Tensor:
    """models/clip/text_encoder.py:113-134."""
    pos = sd[pfx + "positional_embedding"]
    if text.shape[1] != pos.shape[0]:
        raise ValueError(f"length of input should be {pos.shape[0]} but found {text.shape[1]}")
    x = sd[pfx + "token_embedding.weight"][text] + pos                        # :118-119
    for l in range(_num_layers(sd, pfx)):                                     # :121 (causal)
        x = encoder_layer(x, sd, f"{pfx}encoder.layers.{l}.", heads, causal=True)
    hs = layer_norm(x, sd[pfx + "ln_final.weight"], sd[pfx + "ln_final.bias"])  # :125
    if return_hidden_state:
        return hs
    eot = text.argmax(dim=-1)                                                 # :130-132 (first maximum)
    return hs[torch.arange(hs.shape[0], device=hs.device), eot] @ sd[pfx + "projection.weight"].t()


def normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    """F.normalize(x) as used at models/clip/model.py:72-73: x / max(||x||_2, eps) along dim=1."""
    return x / x.norm(dim=1, keepdim=True).clamp_min(eps)


def clip_forward(image: Tensor, text: Tensor, sd: Dict[str, Tensor], img_heads: int, txt_heads: int) -> Tuple[Tensor, Tensor]:
    """models/clip/model.py:65-74."""
    a = vit_encoder(image, sd, "encoder_a.", img_heads)
    b = text_encoder(text, sd, "encoder_b.", txt_heads)
    return normalize(a), normalize(b)


def cross_entropy(logits: Tensor, labels: Tensor, label_smoothing: float = 0.0) -> Tensor:
    """F.cross_entropy(reduction='mean') incl. label smoothing, from log-softmax."""
    lsm = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    nll = -lsm.gather(1, labels.view(-1, 1)).squeeze(1)
    smooth = -lsm.mean(dim=-1)
    return ((1.0 - label_smoothing) * nll + label_smoothing * smooth).mean()


def contrastive_loss(a: Tensor, b: Tensor, logit_scale: Tensor, a_all: Optional[Tensor] = None,
                     b_all: Optional[Tensor] = None, rank: int = 0, label_smoothing: float = 0.0,
                     mask: Optional[Tensor] = None):
    """modules/losses/contrastive_loss_with_temperature.py:50-115.  a_all / b_all are the (already gathered,
    concatenated in rank order) global embeddings; None means single process (:31-33)."""
    T = torch.exp(logit_scale)                                                # :81
    a_all = a if a_all is None else a_all
    b_all = b if b_all is None else b_all
    labels = a.shape[0] * rank + torch.arange(a.shape[0], device=a.device)    # :39-41
    logits_a = a @ b_all.t() * T                                              # :90-92
    logits_b = b @ a_all.t() * T                                              # :93-95
    if mask is not None:                                                      # :97-100
        logits_a, logits_b, labels = logits_a[mask], logits_b[mask], labels[mask]
    loss_a = cross_entropy(logits_a, labels, label_smoothing)                 # :105
    loss_b = cross_entropy(logits_b, labels, label_smoothing)                 # :106
    return (loss_a + loss_b) / 2, logits_a, logits_b, loss_a, loss_b          # :107-115


def clamp_logit_scale(logit_scale: Tensor, lo: Optional[float] = math.log(1), hi: Optional[float] = math.log(100)) -> Tensor:
    """ContrastiveLossWithTemperature.forward clamps .data in place each call (:193)."""
    return logit_scale.clamp(lo, hi)


# ----------------------------------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY.md §8(d)
# ----------------------------------------------------------------------------------------------------------------
def synthetic_batch(B: int, rank: int = 0, image_size: int = 224, ctx: int = 77, vocab: int = 49408,
                    device: str = "cpu") -> Tuple[Tensor, Tensor]:
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    img = torch.randn(B, 3, image_size, image_size, generator=g)
    txt = torch.randint(1, vocab - 2, (B, ctx), generator=g)
    txt[:, -1] = vocab - 1  # EOT is the unique maximum
    return img.to(device), txt.to(device)


def bf16_round_(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Round float tensors to bf16-representable fp32 (parity runs: removes operand-rounding as an error source)."""
    return {k: (v.bfloat16().float() if v.is_floating_point() else v) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------------------------
# Distributed restatement (needs an initialised torch.distributed group; gloo on CPU or nccl on GPU)
# ----------------------------------------------------------------------------------------------------------------
def gather_tensor(t: Tensor, mode: str):
    """utils/distributed.py:28-58.  mode: 'GLOBAL' (autograd all_gather: backward = reduce_scatter),
    'LOCAL' (no-grad gather, own slot replaced by the live tensor), 'NONE' (no-grad gather)."""
    import torch.distributed as dist
    from torch.distributed.nn.functional import all_gather as all_gather_with_backprop

    if mode == "GLOBAL":
        return list(all_gather_with_backprop(t))
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t.detach())
    if mode == "LOCAL":
        out[dist.get_rank()] = t
    return out


def contrastive_loss_distributed(a: Tensor, b: Tensor, logit_scale: Tensor, mode: str = "GLOBAL",
                                 label_smoothing: float = 0.0, mask: Optional[Tensor] = None):
    """contrastive_loss_with_temperature.py:26-115 with torch.distributed initialised (:35-45)."""
    import torch.distributed as dist

    a_all = torch.cat(gather_tensor(a, mode))
    b_all = torch.cat(gather_tensor(b, mode))
    return contrastive_loss(a, b, logit_scale, a_all, b_all, dist.get_rank(), label_smoothing, mask)


def contrastive_grads_lse_exchange(a: Tensor, b: Tensor, logit_scale: Tensor, mode: str = "GLOBAL",
                                   label_smoothing: float = 0.0, mask: Optional[Tensor] = None):
    """Restatement of the CUDA schedule (multimodal_b200/engine_loss.py): no gradient traffic — each rank rebuilds
    d(sum over ranks of loss)/d(its embeddings) from its own logits row block and the peers' row-LSE vectors.
    Returns (loss, dA, dB, dlogit_scale) for THIS rank; must equal autograd over contrastive_loss_distributed."""
    import torch.distributed as dist

    W, r = dist.get_world_size(), dist.get_rank()
    B = a.shape[0]
    N = W * B
    eps = label_smoothing

    def gather(t):
        out = [torch.zeros_like(t) for _ in range(W)]
        dist.all_gather(out, t.contiguous())
        return torch.cat(out)

    with torch.no_grad():
        T = torch.exp(logit_scale)
        a_all, b_all = gather(a), gather(b)
        La, Lb = a @ b_all.t() * T, b @ a_all.t() * T
        lse_a, lse_b = torch.logsumexp(La, 1), torch.logsumexp(Lb, 1)
        lse_a_all, lse_b_all = gather(lse_a), gather(lse_b)
        y = torch.zeros(B, N, dtype=a.dtype, device=a.device)
        y[torch.arange(B), r * B + torch.arange(B)] = 1.0
        t = (1 - eps) * y + eps / N
        # per-row weights of this rank's mean (mask_i / count) and, gathered, of every global row in ITS rank's mean
        w = torch.full((B,), 1.0 / B, dtype=a.dtype) if mask is None else mask.to(a.dtype) / mask.sum().to(a.dtype)
        w_all = gather(w)
        own_a, own_b = torch.exp(La - lse_a[:, None]) - t, torch.exp(Lb - lse_b[:, None]) - t
        col = torch.zeros(N, dtype=torch.bool, device=a.device)
        if mode == "GLOBAL":
            col[:] = True
        elif mode == "LOCAL":
            col[r * B:(r + 1) * B] = True
        tr_a = (torch.exp(La - lse_b_all[None, :]) - t) * (col * w_all)[None, :]
        tr_b = (torch.exp(Lb - lse_a_all[None, :]) - t) * (col * w_all)[None, :]
        dA = 0.5 * T * (own_a * w[:, None] + tr_a) @ b_all
        dB = 0.5 * T * (own_b * w[:, None] + tr_b) @ a_all
        dS = 0.5 * ((own_a * La * w[:, None]).sum() + (own_b * Lb * w[:, None]).sum())
        nll_a = lse_a - La[torch.arange(B), r * B + torch.arange(B)]
        nll_b = lse_b - Lb[torch.arange(B), r * B + torch.arange(B)]
        loss = 0.5 * ((((1 - eps) * nll_a + eps * (lse_a - La.mean(1))) * w).sum()
                      + (((1 - eps) * nll_b + eps * (lse_b - Lb.mean(1))) * w).sum())
    return loss, dA, dB, dS


# ----------------------------------------------------------------------------------------------------------------
# Timing variant (bench.py cpu_baseline / --impl reference): the same restatement expressed with the fused library
# calls the reference itself dispatches to on CPU (F.linear / F.layer_norm / F.scaled_dot_product_attention), so the
# CPU arm is not handicapped by the op-by-op form above.  Checked against clip_forward in tests/test_oracle_cpu.py.
# ----------------------------------------------------------------------------------------------------------------
def _encoder_layer_fused(x: Tensor, sd: Dict[str, Tensor], pfx: str, heads: int, causal: bool) -> Tensor:
    import torch.nn.functional as F

    B, S, d = x.shape
    h = F.layer_norm(x, (d,), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], 1e-5)
    qkv = F.linear(h, sd[pfx + "self_attn.in_proj_weight"], sd[pfx + "self_attn.in_proj_bias"])
    q, k, v = (t.view(B, S, heads, d // heads).transpose(1, 2) for t in qkv.split(d, dim=-1))
    o = F.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(B, S, d)
    x = x + F.linear(o, sd[pfx + "self_attn.out_proj.weight"], sd[pfx + "self_attn.out_proj.bias"])
    h = F.layer_norm(x, (d,), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], 1e-5)
    h = F.linear(h, sd[pfx + "linear1.weight"], sd[pfx + "linear1.bias"])
    h = torch.sigmoid(1.702 * h) * h
    return x + F.linear(h, sd[pfx + "linear2.weight"], sd[pfx + "linear2.bias"])


def clip_forward_fused(image: Tensor, text: Tensor, sd: Dict[str, Tensor], img_heads: int, txt_heads: int):
    import torch.nn.functional as F

    pfx = "encoder_a."
    w = sd[pfx + "conv.weight"]
    x = F.conv2d(image, w, stride=w.shape[-1]).flatten(2).permute(0, 2, 1)
    x = torch.cat([sd[pfx + "cls_token_embedding"].expand(x.shape[0], 1, -1), x], dim=1) + sd[pfx + "positional_embedding"]
    x = F.layer_norm(x, x.shape[-1:], sd[pfx + "ln_pre.weight"], sd[pfx + "ln_pre.bias"], 1e-5)
    for l in range(_num_layers(sd, pfx)):
        x = _encoder_layer_fused(x, sd, f"{pfx}encoder.layers.{l}.", img_heads, False)
    a = F.layer_norm(x[:, 0, :], x.shape[-1:], sd[pfx + "ln_post.weight"], sd[pfx + "ln_post.bias"], 1e-5) @ sd[pfx + "projection"]
    pfx = "encoder_b."
    y = F.embedding(text, sd[pfx + "token_embedding.weight"]) + sd[pfx + "positional_embedding"]
    for l in range(_num_layers(sd, pfx)):
        y = _encoder_layer_fused(y, sd, f"{pfx}encoder.layers.{l}.", txt_heads, True)
    y = F.layer_norm(y, y.shape[-1:], sd[pfx + "ln_final.weight"], sd[pfx + "ln_final.bias"], 1e-5)
    b = F.linear(y[torch.arange(y.shape[0]), text.argmax(dim=-1)], sd[pfx + "projection.weight"])
    return F.normalize(a), F.normalize(b)
