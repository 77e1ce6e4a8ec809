"""Standalone (single-module) forward runtimes for the layer types of the hot path, so that the modules the reference
exposes — and tests — on their own are callable outside an encoder:

  modules/layers/multi_head_attention.py:19-80   MultiHeadSelfAttention
  modules/layers/transformer.py:31-154           TransformerEncoderLayer (pre- and post-norm), :157-259 TransformerEncoder
  modules/layers/patch_embedding.py:25-157       PatchEmbeddings
  modules/layers/mlp.py:13-66                    MLP ([Linear, activation, Linear] form)

They launch exactly the kernels the fused encoder schedules launch (tcgen05 GEMMs with bias / activation epilogues,
tcgen05 attention, the add+LayerNorm kernel); nothing is computed by PyTorch.  Pre-norm `TransformerEncoderLayer` /
`TransformerEncoder` also train on their own (grad mode on: `engine_coca_train.LayersTrainRuntime`, the CLIP towers' fused
forward / backward schedule); the other standalone modules compute forward values only, and asking them for an autograd
graph raises instead of returning detached tensors.
Shape limits are those of the kernels: head_dim 64 (fused attention; 96 / 128 and arbitrary boolean masks go through
the general kernel), feature sizes multiples of 8, 3-channel images.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import ops
from ._lib import MMBError
from .engine import Workspace
from .engine_flava import _Shadows


def forward_only_guard(mod: nn.Module, what: str) -> None:
    if torch.is_grad_enabled() and any(p.requires_grad for p in mod.parameters()):
        raise MMBError(f"{what} (multimodal_b200) computes forward values only when called on its own — its backward "
                       "exists inside the CLIP towers' fused schedule.  Call it under torch.no_grad().")


def _wants_graph(mod: nn.Module, x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in mod.parameters()))


def _cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise MMBError(f"{what}: expected a CUDA tensor (multimodal_b200 has no CPU path), got device {t.device}")


class _Rt:
    """Per-module scratch: reused workspace + bf16 weight shadows (re-cast when a parameter changes)."""

    def __init__(self, device):
        self.ws, self.sh, self.device = Workspace(device), _Shadows(device), device


def _rt(mod: nn.Module, device) -> _Rt:
    rt = getattr(mod, "_mmb_rt", None)
    if rt is None or rt.device != device:
        rt = _Rt(device)
        object.__setattr__(mod, "_mmb_rt", rt)   # not a submodule / buffer: keeps the state dict untouched
    return rt


def _bool_mask_u8(mask: Optional[torch.Tensor], B: int, S: int, what: str) -> Optional[torch.Tensor]:
    """bool [B, S, S] or [B, 1, S, S] (True = attend) -> uint8 [B, S, S]; float (additive) masks are not supported."""
    if mask is None:
        return None
    if mask.dtype != torch.bool:
        raise NotImplementedError(f"{what}: only boolean attention masks (True = attend) are on the accelerated path")
    if mask.dim() == 4:
        if mask.shape[1] != 1:
            raise NotImplementedError(f"{what}: per-head masks are not on the accelerated path (the head dim must be 1)")
        mask = mask[:, 0]
    if mask.dim() == 2:
        mask = mask[None].expand(B, S, S)
    if tuple(mask.shape) != (B, S, S):
        raise ValueError(f"{what}: attention mask shape {tuple(mask.shape)} does not match [{B}, {S}, {S}]")
    return mask.to(torch.uint8).contiguous()


def _attention(rt: _Rt, QKV: torch.Tensor, O: torch.Tensor, B: int, S: int, H: int, hd: int, mask_u8, causal: bool):
    d = H * hd
    scale = 1.0 / math.sqrt(hd)
    if mask_u8 is None and hd == 64 and S <= 384:
        ops.attention_fwd(QKV, O, None, B, S, H, causal, scale)
    else:
        ops.attention_fwd_generic(QKV[:, :d], QKV[:, d:2 * d], QKV[:, 2 * d:], O, B=B, Sq=S, Skv=S, H=H, head_dim=hd,
                                  bsq=S * 3 * d, bsk=S * 3 * d, bsv=S * 3 * d, bso=S * d, scale=scale, mask=mask_u8,
                                  mask_bs=S * S if mask_u8 is not None else 0, mask_qs=S if mask_u8 is not None else 0,
                                  causal=causal)


def _to_bf16_rows(rt: _Rt, x: torch.Tensor, name: str) -> torch.Tensor:
    xf = x.contiguous().float()
    out = rt.ws.get(name, (xf.numel() // xf.shape[-1], xf.shape[-1]), torch.bfloat16)
    ops.cast_bf16(xf.view(-1), out.view(-1))
    return out


# ---------------------------------------------------------------------------------------------------------------------
def mhsa_forward(mod: nn.Module, query: torch.Tensor, attn_mask: Optional[torch.Tensor] = None,
                 is_causal: bool = False) -> torch.Tensor:
    """MultiHeadSelfAttention.forward (multi_head_attention.py:39-80): input_proj -> SDPA -> output_proj."""
    forward_only_guard(mod, "MultiHeadSelfAttention")
    _cuda(query, "MultiHeadSelfAttention")
    B, S, d = query.shape
    H = mod.num_heads
    if d % H or (d // H) not in (64, 96, 128) or d % 8:
        raise MMBError(f"MultiHeadSelfAttention: head_dim {d / H:g} is not supported by the attention kernels (64 / 96 / 128)")
    rt = _rt(mod, query.device)
    bf = torch.bfloat16
    Xb = _to_bf16_rows(rt, query, "mhsa.X")
    QKV = rt.ws.get("mhsa.QKV", (B * S, 3 * d), bf)
    O = rt.ws.get("mhsa.O", (B * S, d), bf)
    ops.gemm(Xb, rt.sh.get("wqkv", [mod.input_proj.weight]), bias=mod.input_proj.bias, out=QKV)
    _attention(rt, QKV, O, B, S, H, d // H, _bool_mask_u8(attn_mask, B, S, "MultiHeadSelfAttention"), bool(is_causal))
    out = torch.empty((B * S, d), device=query.device, dtype=torch.float32)
    ops.gemm(O, rt.sh.get("wo", [mod.output_proj.weight]), bias=mod.output_proj.bias, epilogue=ops.EPI_F32, out=out)
    return out.view(B, S, d).to(query.dtype)


def _act_code(act: nn.Module) -> int:
    from .modules.layers.activation import SiLU

    if isinstance(act, nn.GELU) and getattr(act, "approximate", "none") == "none":
        return ops.ACT_GELU_ERF
    if isinstance(act, SiLU):
        return ops.ACT_QUICK_GELU
    raise MMBError(f"unsupported MLP activation {type(act).__name__} on the accelerated path (nn.GELU / SiLU)")


def mlp_forward(mod: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """MLP.forward (mlp.py:62-66) for the transformer form [Linear, activation, (Dropout 0), Linear]."""
    forward_only_guard(mod, "MLP")
    _cuda(x, "MLP")
    seq = [m for m in mod.model if not isinstance(m, nn.Dropout)]
    if len(seq) != 3 or not isinstance(seq[0], nn.Linear) or not isinstance(seq[2], nn.Linear):
        raise MMBError("standalone MLP supports the [Linear, activation, Linear] form (one hidden layer, no normalisation)")
    rt = _rt(mod, x.device)
    act = _act_code(seq[1])
    Xb = _to_bf16_rows(rt, x, "mlp.X")
    M = Xb.shape[0]
    ff, dout = seq[0].weight.shape[0], seq[2].weight.shape[0]
    PRE = rt.ws.get("mlp.PRE", (M, ff), torch.bfloat16)
    HACT = rt.ws.get("mlp.HACT", (M, ff), torch.bfloat16)
    ops.gemm(Xb, rt.sh.get("w1", [seq[0].weight]), bias=seq[0].bias, epilogue=ops.EPI_BF16_ACT, out=PRE, out2=HACT, act=act)
    out = torch.empty((M, dout), device=x.device, dtype=torch.float32)
    ops.gemm(HACT, rt.sh.get("w2", [seq[2].weight]), bias=seq[2].bias, epilogue=ops.EPI_F32, out=out)
    return out.view(*x.shape[:-1], dout).to(x.dtype)


def encoder_layer_forward(mod: nn.Module, hidden_states: torch.Tensor,
                          attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """TransformerEncoderLayer.forward (transformer.py:95-154), pre-norm (:95-111) and post-norm (:113-128)."""
    if _wants_graph(mod, hidden_states):
        if not mod.norm_first:
            raise MMBError("standalone post-norm TransformerEncoderLayer has no backward schedule; call it under "
                           "torch.no_grad()")
        from .engine_coca_train import standalone_layers
        B, S, _ = hidden_states.shape
        return standalone_layers(mod, [mod], None, hidden_states,
                                 _bool_mask_u8(attention_mask, B, S, "TransformerEncoderLayer"))[0]
    _cuda(hidden_states, "TransformerEncoderLayer")
    B, S, d = hidden_states.shape
    at, mlp = mod.attention, mod.feedforward.model
    H = at.num_heads
    hd = d // H
    if d % H or hd not in (64, 96, 128):
        raise MMBError(f"TransformerEncoderLayer: head_dim {d / H:g} is not supported by the attention kernels")
    rt = _rt(mod, hidden_states.device)
    ws, sh = rt.ws, rt.sh
    M = B * S
    bf, f32 = torch.bfloat16, torch.float32
    act = _act_code(mlp[1])
    ff = mlp[0].weight.shape[0]
    mask_u8 = _bool_mask_u8(attention_mask, B, S, "TransformerEncoderLayer")
    X = hidden_states.contiguous().float().view(M, d)
    LN, QKV, O = ws.get("l.LN", (M, d), bf), ws.get("l.QKV", (M, 3 * d), bf), ws.get("l.O", (M, d), bf)
    Y, PRE, HACT = ws.get("l.Y", (M, d), bf), ws.get("l.PRE", (M, ff), bf), ws.get("l.HACT", (M, ff), bf)
    ln1, ln2 = mod.attention_layernorm, mod.feedforward_layernorm
    wqkv, wo = sh.get("wqkv", [at.input_proj.weight]), sh.get("wo", [at.output_proj.weight])
    w1, w2 = sh.get("w1", [mlp[0].weight]), sh.get("w2", [mlp[-1].weight])
    out = torch.empty((M, d), device=X.device, dtype=f32)
    if mod.norm_first:
        ops.add_layernorm_fwd(X, None, None, LN, None, ln1.weight, ln1.bias, None, None, M, d, ln1.eps)
        ops.gemm(LN, wqkv, bias=at.input_proj.bias, out=QKV)
        _attention(rt, QKV, O, B, S, H, hd, mask_u8, False)
        ops.gemm(O, wo, bias=at.output_proj.bias, out=Y)
        XM = ws.get("l.XM", (M, d), f32)                       # x + attention(LN(x)); LN2 of it for the MLP
        ops.add_layernorm_fwd(X, Y, XM, LN, None, ln2.weight, ln2.bias, None, None, M, d, ln2.eps)
        ops.gemm(LN, w1, bias=mlp[0].bias, epilogue=ops.EPI_BF16_ACT, out=PRE, out2=HACT, act=act)
        ops.gemm(HACT, w2, bias=mlp[-1].bias, out=Y)
        # out = XM + mlp: the add kernel with an identity-free LayerNorm is not needed — reuse add+LN writing only x_out
        ops.add_layernorm_fwd(XM, Y, out, LN, None, ln2.weight, ln2.bias, None, None, M, d, ln2.eps)
    else:
        ops.cast_bf16(X.view(-1), LN.view(-1))                 # attention(x) on the raw input
        ops.gemm(LN, wqkv, bias=at.input_proj.bias, out=QKV)
        _attention(rt, QKV, O, B, S, H, hd, mask_u8, False)
        ops.gemm(O, wo, bias=at.output_proj.bias, out=Y)
        H1 = ws.get("l.H1", (M, d), f32)                       # LN1(x + attention(x)), fp32 + its bf16 operand copy
        ops.add_layernorm_fwd(X, Y, None, LN, H1, ln1.weight, ln1.bias, None, None, M, d, ln1.eps)
        ops.gemm(LN, w1, bias=mlp[0].bias, epilogue=ops.EPI_BF16_ACT, out=PRE, out2=HACT, act=act)
        ops.gemm(HACT, w2, bias=mlp[-1].bias, out=Y)
        ops.add_layernorm_fwd(H1, Y, None, None, out, ln2.weight, ln2.bias, None, None, M, d, ln2.eps)
    return out.view(B, S, d).to(hidden_states.dtype)


def encoder_forward(mod: nn.Module, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                    return_hidden_states: bool = False):
    """TransformerEncoder.forward (transformer.py:216-259): the layers (+ optional final LayerNorm)."""
    from .modules.layers.transformer import TransformerOutput

    if _wants_graph(mod, hidden_states):
        if not all(layer.norm_first for layer in mod.layer):
            raise MMBError("standalone post-norm TransformerEncoder has no backward schedule; call it under torch.no_grad()")
        from .engine_coca_train import standalone_layers
        B, S, _ = hidden_states.shape
        out, hidden = standalone_layers(mod, mod.layer, mod.final_layer_norm, hidden_states,
                                        _bool_mask_u8(attention_mask, B, S, "TransformerEncoder"))
        return TransformerOutput(last_hidden_state=out, hidden_states=hidden if return_hidden_states else None)
    _cuda(hidden_states, "TransformerEncoder")
    x = hidden_states
    all_hidden = [x] if return_hidden_states else None
    with torch.no_grad():
        for layer in mod.layer:
            x = encoder_layer_forward(layer, x, attention_mask)
            if return_hidden_states:
                all_hidden.append(x)
        if mod.final_layer_norm is not None:
            x = mod.final_layer_norm(x)
    return TransformerOutput(last_hidden_state=x, hidden_states=all_hidden)


def patch_embeddings_forward(mod: nn.Module, image: torch.Tensor, image_patches_mask: Optional[torch.Tensor] = None):
    """PatchEmbeddings.forward (patch_embedding.py:104-154) without random patch dropping: conv projection as
    im2col + tcgen05 GEMM, [cls |] patches (mask-token substitution) + position embeddings in one assembly kernel."""
    from .modules.layers.patch_embedding import PatchEmbeddingsOutput

    forward_only_guard(mod, "PatchEmbeddings")
    _cuda(image, "PatchEmbeddings")
    conv = mod.conv_projection
    d, ps = conv.weight.shape[0], conv.weight.shape[2]
    img = image.contiguous().float()
    B, C, Hh, Ww = img.shape
    if C != 3 or (Hh, Ww) != tuple(mod.image_size):
        raise ValueError(f"Input image shape {tuple(image.shape)} doesn't match the model's 3 x {mod.image_size}")
    P = (Hh // ps) * (Ww // ps)
    S = P + (1 if mod.include_cls_embed else 0)
    K = 3 * ps * ps
    Kp = -(-K // 8) * 8
    rt = _rt(mod, image.device)
    ws, sh = rt.ws, rt.sh
    bf = torch.bfloat16
    PATCH = ws.get("pe.PATCH", (B * P, Kp), bf)[:, :K]
    PO = ws.get("pe.PO", (B * P, d), bf)
    ops.im2col(img, ps, PATCH)
    w = sh.get("conv.w", [conv.weight.view(d, K)])
    if Kp != K:
        wp = ws.get("pe.WCONV", (d, Kp), bf)[:, :K]
        wp.copy_(w)
        w = wp
    ops.gemm(PATCH, w, bias=conv.bias, out=PO)
    pm = None
    if image_patches_mask is not None and mod.mask_token is not None:
        pm = image_patches_mask.reshape(B, P).to(torch.uint8).contiguous()
    X = torch.empty((B * S, d), device=image.device, dtype=torch.float32)
    ops.vit_assemble_fwd(PO, mod.cls_token if mod.include_cls_embed else None, mod.position_embeddings,
                         mod.mask_token if pm is not None else None, pm, X, B, S, d)
    return PatchEmbeddingsOutput(embeddings=X.view(B, S, d).to(image.dtype))
