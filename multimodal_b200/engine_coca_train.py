"""Training runtime of the CoCa model family (BASELINE.json config 5 / SURVEY.md §8 a14 + f3 as a training step):
forwards that keep what the backward needs + explicit backward schedules, behind torch.autograd Functions, so that
`CoCaModel` / `CoCaForPretraining` train with ``loss.backward()`` like the reference (models/coca/coca_model.py:69-130,
398-454 under autograd).

All layer stacks run on ``engine.TransformerStack`` (the CLIP towers' fused schedule and kernels):
  vision encoder      : packed `input_proj` self-attention (tcgen05 fwd / bwd), erf-GELU MLP, optional final LayerNorm
                        (modules/encoders/vision_transformer.py:56-89, patch_embedding.py:104-154)
  text decoder        : separate q / k / v projections presented as one packed operand, the [B, S, S] causal x padding
                        mask on the general attention kernels (fwd: mma.sync, bwd: SIMT), CLS row -> ln_final -> projection
                        (models/coca/text_decoder.py:141-203)
  multimodal decoder  : causal self-attention (tcgen05) + cross-attention to the pooled image tokens (general kernels) +
                        MLP per layer, final LayerNorm (models/coca/multimodal_decoder.py:86-108)
  attention pooler    : LayerNorm-ed keys / values, batch-shared learned queries (their gradient is summed over the batch
                        with fp32 atomics), ln_post (modules/layers/attention_pooler.py:48-72)
  vocabulary head     : Linear -> CrossEntropy(ignore_index) with materialised fp32 logits in training (the forward-only
                        path keeps the fused statistics GEMM), backward = d logits kernel + two GEMMs
Every training forward keeps its activations in its own Workspace (held by the autograd node).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops
from ._lib import MMBError
from .engine import ParamStore, TransformerStack, Workspace
from .engine_flava_train import wants_grad  # noqa: F401  (re-exported for the modules)


def _act_code(act: nn.Module) -> int:
    if isinstance(act, nn.GELU) and getattr(act, "approximate", "none") == "none":
        return ops.ACT_GELU_ERF
    raise MMBError(f"unsupported MLP activation {type(act).__name__} on the accelerated path (nn.GELU only)")


def _qkv_first(layers) -> List[nn.Parameter]:
    """Parameter order that makes the separate q / k / v (and cross k / v) projections packable."""
    out: List[nn.Parameter] = []
    for layer in layers:
        at = layer.attention
        if not hasattr(at, "input_proj"):
            out += [at.q_proj.weight, at.k_proj.weight, at.v_proj.weight, at.q_proj.bias, at.k_proj.bias, at.v_proj.bias]
        ca = getattr(layer, "cross_attention", None)
        if ca is not None:
            out += [ca.k_proj.weight, ca.v_proj.weight, ca.k_proj.bias, ca.v_proj.bias]
    return out


def _store_for(owner: nn.Module, first: Sequence[nn.Parameter]) -> ParamStore:
    params, seen = list(first), {id(p) for p in first}
    for p in owner.parameters():
        if id(p) not in seen:
            seen.add(id(p))
            params.append(p)
    return ParamStore(params)


def _adapters(layers, st: ParamStore):
    out = []
    for layer in layers:
        at, mlp = layer.attention, layer.feedforward.model
        if hasattr(at, "input_proj"):     # MultiHeadSelfAttention: already packed [3d, d]
            attn = SimpleNamespace(in_proj_weight=at.input_proj.weight, in_proj_bias=at.input_proj.bias,
                                   out_proj=at.output_proj, num_heads=at.num_heads)
        else:
            attn = SimpleNamespace(in_proj_weight=st.pack([at.q_proj.weight, at.k_proj.weight, at.v_proj.weight]),
                                   in_proj_bias=st.pack([at.q_proj.bias, at.k_proj.bias, at.v_proj.bias], fp32=True),
                                   out_proj=at.output_proj, num_heads=at.num_heads)
        ad = SimpleNamespace(self_attn=attn, norm1=layer.attention_layernorm, norm2=layer.feedforward_layernorm,
                             linear1=mlp[0], linear2=mlp[-1])
        ca = getattr(layer, "cross_attention", None)
        if ca is not None:
            ad.cross_attn = SimpleNamespace(q_w=ca.q_proj.weight, q_b=ca.q_proj.bias,
                                            kv_w=st.pack([ca.k_proj.weight, ca.v_proj.weight]),
                                            kv_b=st.pack([ca.k_proj.bias, ca.v_proj.bias], fp32=True),
                                            out_proj=ca.output_proj)
            ad.norm_cross = layer.cross_attention_layernorm
        out.append(ad)
    return out


class _Stack:
    """ParamStore + TransformerStack over a list of TorchMultimodal pre-norm encoder / decoder layers."""

    def __init__(self, owner: nn.Module, layers, prefix: str, causal: bool):
        layers = list(layers)
        l0 = layers[0]
        if not l0.norm_first:
            raise MMBError("only pre-norm (norm_first=True) layers are on the accelerated path")
        self.store = _store_for(owner, _qkv_first(layers))
        self.device = self.store.device
        self.d = l0.attention_layernorm.normalized_shape[0]
        H = l0.attention.num_heads
        if self.d // H != 64:
            raise MMBError(f"training needs head_dim 64 in the layer stacks (got {self.d // H}); the poolers may differ")
        self.ws = Workspace(self.device)
        self.stack = TransformerStack(_adapters(layers, self.store), self.store, self.ws, d=self.d, heads=H,
                                      ff=l0.feedforward.model[0].weight.shape[0], causal=causal,
                                      act=_act_code(l0.feedforward.model[1]), prefix=prefix)
        self.prefix, self.L = prefix, len(layers)

    def finish(self, XM, Y, M: int, ln: Optional[nn.Module], save: Workspace):
        """XF = XM + Y (the residual stream after the last layer); LAST = ln(XF) when a final LayerNorm exists."""
        d, pfx = self.d, self.prefix
        f32 = torch.float32
        XF = torch.empty((M, d), device=self.device, dtype=f32)
        LAST = torch.empty((M, d), device=self.device, dtype=f32) if ln is not None else None
        aff = ln if ln is not None else self.stack.layers[0].norm1   # affine terms unused when nothing is normalised
        ops.add_layernorm_fwd(XM, Y, XF, None, LAST, aff.weight, aff.bias,
                              save.get(f"{pfx}.mF", (M,), f32) if ln is not None else None,
                              save.get(f"{pfx}.rF", (M,), f32) if ln is not None else None, M, d, aff.eps)
        save.XF = XF
        return XF, LAST

    def start_backward(self, save: Workspace, M: int, ln: Optional[nn.Module], dLAST, dXF):
        """(G fp32, Gb bf16, top_bias_done): gradient w.r.t. XF entering the stack's backward."""
        d, pfx, st = self.d, self.prefix, self.store
        f32, bf = torch.float32, torch.bfloat16
        G = self.ws.get(f"{pfx}.G", (M, d), f32)
        Gb = self.ws.get(f"{pfx}.Gb", (M, d), bf)
        if ln is not None and dLAST is not None:
            ops.layernorm_bwd(save.XF, None, dLAST, save.get(f"{pfx}.mF", (M,), f32), save.get(f"{pfx}.rF", (M,), f32),
                              ln.weight, dXF, G, Gb, st.grad(ln.weight), st.grad(ln.bias), M, d,
                              gsum=self.stack.top_bias_grad())
            return G, Gb, True
        if dXF is None:
            ops.zero_(G)
        else:
            G.copy_(dXF.view(M, d))      # the stack's backward works in place on G
        ops.cast_bf16(G, Gb)
        return G, Gb, False


def _f32(t: Optional[torch.Tensor], shape) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous().float().view(shape)


# ---------------------------------------------------------------------------------------------------------------------
class VisionTrainRuntime:
    """VisionTransformer (modules/encoders/vision_transformer.py:56-89).  Output: last_hidden_state [B*S, d]."""

    def __init__(self, mod: nn.Module):
        self.mod = mod
        self.s = _Stack(mod, mod.encoder.layer, "cvit", causal=False)
        self.store = self.s.store

    def forward(self, data, diff):
        images, image_patches_mask = data
        emb, s, st = self.mod.embeddings, self.s, self.store
        d = s.d
        conv = emb.conv_projection
        ps = conv.weight.shape[2]
        image = images.contiguous().float()
        B, _, Hh, Ww = image.shape
        P = (Hh // ps) * (Ww // ps)
        cls = emb.cls_token if emb.include_cls_embed else None
        S = P + (1 if cls is not None else 0)
        K = 3 * ps * ps
        Kp = -(-K // 8) * 8
        bf, f32 = torch.bfloat16, torch.float32
        st.refresh()
        save = Workspace(s.device)
        PATCH = save.get("cvit.PATCH", (B * P, Kp), bf)[:, :K]
        PO = s.ws.get("cvit.PO", (B * P, d), bf)
        X0 = torch.empty((B * S, d), device=image.device, dtype=f32)
        ops.im2col(image, ps, PATCH)
        w = st.shadow2d(conv.weight)
        if Kp != K:
            wp = s.ws.get("cvit.WCONV", (d, Kp), bf)[:, :K]
            wp.copy_(w)
            w = wp
        ops.gemm(PATCH, w, bias=conv.bias, out=PO)
        pm = None
        if image_patches_mask is not None and emb.mask_token is not None:
            pm = image_patches_mask.reshape(B, P).to(torch.uint8).contiguous()
        ops.vit_assemble_fwd(PO, cls, emb.position_embeddings, emb.mask_token if pm is not None else None, pm, X0, B, S, d)
        XM, Y = s.stack.forward(X0, B, S, True, save=save)
        XF, LAST = s.finish(XM, Y, B * S, self.mod.encoder.final_layer_norm, save)
        save.B, save.S, save.P, save.K, save.pm = B, S, P, K, pm
        self.last_hidden = ([X0.view(B, S, d)] + [save.bufs[f"cvit.XA.{l}"].view(B, S, d) for l in range(1, s.L)]
                            + [XF.view(B, S, d)])
        return ((LAST if LAST is not None else XF),), save

    def backward(self, save, dOUT):
        emb, s, st = self.mod.embeddings, self.s, self.store
        d, B, S, P, K = s.d, save.B, save.S, save.P, save.K
        M = B * S
        fln = self.mod.encoder.final_layer_norm
        dOUT = _f32(dOUT, (M, d))
        G, Gb, done = s.start_backward(save, M, fln, dOUT if fln is not None else None, None if fln is not None else dOUT)
        G = s.stack.backward(G, Gb, B, S, top_bias_done=done, save=save)
        has_cls = emb.include_cls_embed
        conv = emb.conv_projection
        ops.batch_sum(G, st.grad(emb.position_embeddings), B, S * d, S * d)
        if has_cls:
            ops.batch_sum(G, st.grad(emb.cls_token), B, S * d, d)
        DP = s.ws.get("cvit.DP", (B * P, d), torch.bfloat16)
        ops.vit_assemble_bwd(G, save.pm, DP, st.grad(emb.mask_token) if save.pm is not None else None, B, S, d, has_cls)
        PATCH = save.get("cvit.PATCH", (B * P, -(-K // 8) * 8), torch.bfloat16)[:, :K]
        ops.gemm(DP, PATCH, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad2d(conv.weight),
                 splits=ops.wgrad_splits(d, PATCH.shape[1], B * P), accumulate=True)
        if conv.bias is not None:
            ops.colsum_bf16(DP, st.grad(conv.bias), B * P, d, d)
        return ()


class PoolerTrainRuntime:
    """AttentionPooler (modules/layers/attention_pooler.py:16-72).  Input x [B, S, d_in] (differentiable)."""

    def __init__(self, mod: nn.Module):
        self.mod = mod
        at = mod.attn
        self.store = _store_for(mod, [at.k_proj.weight, at.v_proj.weight, at.k_proj.bias, at.v_proj.bias])
        st = self.store
        self.kv_w = st.pack([at.k_proj.weight, at.v_proj.weight])
        self.kv_b = st.pack([at.k_proj.bias, at.v_proj.bias], fp32=True)
        self.ws = Workspace(st.device)
        self.device = st.device

    def _dims(self):
        m = self.mod
        nq, dout = m.query.shape
        H = m.attn.num_heads
        return nq, dout, H, dout // H

    def forward(self, data, diff):
        (x,) = diff
        m, st = self.mod, self.store
        at = m.attn
        B, S, din = x.shape
        nq, dout, H, hd = self._dims()
        if hd not in (64, 96, 128):
            raise MMBError(f"unsupported pooler head_dim {hd}")
        bf, f32 = torch.bfloat16, torch.float32
        st.refresh()
        save = Workspace(self.device)
        x32 = x.contiguous().float().view(B * S, din)
        xk = save.get("xk", (B * S, din), bf)
        ops.add_layernorm_fwd(x32, None, None, xk, None, m.ln_k.weight, m.ln_k.bias, save.get("mk", (B * S,), f32),
                              save.get("rk", (B * S,), f32), B * S, din, m.ln_k.eps)
        qn = save.get("qn", (nq, dout), bf)
        ops.add_layernorm_fwd(m.query.data, None, None, qn, None, m.ln_q.weight, m.ln_q.bias, save.get("mq", (nq,), f32),
                              save.get("rq", (nq,), f32), nq, dout, m.ln_q.eps)
        Qp = save.get("Qp", (nq, dout), bf)
        ops.gemm(qn, st.shadow(at.q_proj.weight), bias=at.q_proj.bias, out=Qp)
        KV = save.get("KV", (B * S, 2 * dout), bf)
        ops.gemm(xk, st.shadow(self.kv_w), bias=self.kv_b, out=KV)
        O = save.get("O", (B * nq, dout), bf)
        ops.attention_fwd_generic(Qp, KV[:, :dout], KV[:, dout:], O, B=B, Sq=nq, Skv=S, H=H, head_dim=hd, bsq=0,
                                  bsk=S * 2 * dout, bsv=S * 2 * dout, bso=nq * dout, scale=1.0 / math.sqrt(hd))
        Y = self.ws.get("Y", (B * nq, dout), bf)
        ops.gemm(O, st.shadow(at.output_proj.weight), bias=at.output_proj.bias, out=Y)
        Y32 = save.get("Y32", (B * nq, dout), f32)
        out = torch.empty((B * nq, dout), device=self.device, dtype=f32)
        ops.add_layernorm_fwd(None, Y, Y32, None, out, m.ln_post.weight, m.ln_post.bias, save.get("mp", (B * nq,), f32),
                              save.get("rp", (B * nq,), f32), B * nq, dout, m.ln_post.eps)
        save.x32, save.B, save.S, save.din = x32, B, S, din
        return (out,), save

    def backward(self, save, dOUT):
        m, st = self.mod, self.store
        at = m.attn
        B, S, din = save.B, save.S, save.din
        nq, dout, H, hd = self._dims()
        bf, f32 = torch.bfloat16, torch.float32
        n = B * nq
        g = lambda name, shape, dt: save.get(name, shape, dt)  # noqa: E731
        dYb = self.ws.get("dYb", (n, dout), bf)
        ops.layernorm_bwd(g("Y32", (n, dout), f32), None, _f32(dOUT, (n, dout)), g("mp", (n,), f32), g("rp", (n,), f32),
                          m.ln_post.weight, None, None, dYb, st.grad(m.ln_post.weight), st.grad(m.ln_post.bias), n, dout,
                          gsum=st.grad(at.output_proj.bias))
        O, KV = g("O", (n, dout), bf), g("KV", (B * S, 2 * dout), bf)
        Qp, qn, xk = g("Qp", (nq, dout), bf), g("qn", (nq, dout), bf), g("xk", (B * S, din), bf)
        ops.gemm(dYb, O, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(at.output_proj.weight),
                 splits=ops.wgrad_splits(dout, dout, n), accumulate=True)
        dO = self.ws.get("dO", (n, dout), bf)
        ops.gemm(dYb, st.shadow(at.output_proj.weight), b_mn=True, out=dO)
        dQ32 = torch.zeros((nq, dout), device=self.device, dtype=f32)   # summed over the batch by the kernel
        dKV = self.ws.get("dKV", (B * S, 2 * dout), bf)
        ops.attention_bwd_generic(Qp, KV[:, :dout], KV[:, dout:], dO, dKV[:, :dout], dKV[:, dout:], dq=None, dq_f32=dQ32,
                                  B=B, Sq=nq, Skv=S, H=H, head_dim=hd, bsq=0, bsk=S * 2 * dout, bsv=S * 2 * dout,
                                  bso=nq * dout, scale=1.0 / math.sqrt(hd))
        dQb = ops.cast_bf16(dQ32)
        ops.gemm(dQb, qn, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(at.q_proj.weight), accumulate=True)
        ops.colsum_bf16(dQb, st.grad(at.q_proj.bias), nq, dout, dout)
        dqn = self.ws.get("dqn", (nq, dout), bf)
        ops.gemm(dQb, st.shadow(at.q_proj.weight), b_mn=True, out=dqn)
        gq = st.grad(m.query)
        ops.layernorm_bwd(m.query.data, dqn, None, g("mq", (nq,), f32), g("rq", (nq,), f32), m.ln_q.weight, gq, gq, None,
                          st.grad(m.ln_q.weight), st.grad(m.ln_q.bias), nq, dout)
        ops.gemm(dKV, xk, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(self.kv_w),
                 splits=ops.wgrad_splits(2 * dout, din, B * S), accumulate=True)
        ops.colsum_bf16(dKV, st.grad(self.kv_b), B * S, 2 * dout, 2 * dout)
        dxk = self.ws.get("dxk", (B * S, din), bf)
        ops.gemm(dKV, st.shadow(self.kv_w), b_mn=True, out=dxk)
        dx = torch.empty((B * S, din), device=self.device, dtype=f32)
        ops.layernorm_bwd(save.x32, dxk, None, g("mk", (B * S,), f32), g("rk", (B * S,), f32), m.ln_k.weight, None, dx,
                          None, st.grad(m.ln_k.weight), st.grad(m.ln_k.bias), B * S, din)
        return (dx.view(B, S, din),)


class TextDecoderTrainRuntime:
    """CoCaTextDecoder with embed_cls=True (models/coca/text_decoder.py:66-203).
    Outputs: pooled [B, out_dim] (projected ln_final(CLS row)) and XF [B*S, d] (tokens = XF[:, :-1])."""

    def __init__(self, mod: nn.Module):
        if not mod.embed_cls:
            raise NotImplementedError("training CoCaTextDecoder(embed_cls=False) is not on the accelerated path")
        if mod.text_projection is None or mod.text_projection.bias is not None:
            raise NotImplementedError("training expects the bias-free text_projection of the reference builder")
        self.mod = mod
        self.s = _Stack(mod, mod.transformer_decoder.layer, "ctxt", causal=True)
        self.store = self.s.store
        self._idx = None

    def forward(self, data, diff):
        input_ids, mask_u8, S = data
        m, s, st = self.mod, self.s, self.store
        d = s.d
        emb = m.embeddings
        ids = input_ids.long().contiguous()
        B = ids.shape[0]
        bf, f32 = torch.bfloat16, torch.float32
        st.refresh()
        save = Workspace(s.device)
        X0 = torch.empty((B * S, d), device=s.device, dtype=f32)
        ops.coca_text_embed_fwd(ids, emb.token_embeddings.weight, emb.cls_embedding, emb.position_embeddings, X0, B, S, d,
                                emb.token_embeddings.weight.shape[0])
        s.stack.causal = mask_u8 is None      # a [B, S, S] mask already contains the causal structure
        XM, Y = s.stack.forward(X0, B, S, True, save=save, mask3=mask_u8)
        XF, _ = s.finish(XM, Y, B * S, None, save)
        if self._idx is None or self._idx.numel() != B:
            self._idx = torch.full((B,), S - 1, dtype=torch.int32, device=s.device)
        ln = getattr(m, "ln_final", None)
        POOLb = save.get("ctxt.POOLb", (B, d), bf)
        if ln is not None:    # LayerNorm of the CLS row only (:186-189): gathered rows
            ops.add_layernorm_fwd(XF, None, save.get("ctxt.XSEL", (B, d), f32), POOLb, None, ln.weight, ln.bias,
                                  save.get("ctxt.mL", (B,), f32), save.get("ctxt.rL", (B,), f32), B, d, ln.eps,
                                  row_idx=self._idx, rows_per_group=S)
        else:
            ops.gather_rows_cast(XF, POOLb, B, S, S - 1, d)
        pooled = torch.empty((B, m.text_projection.weight.shape[0]), device=s.device, dtype=f32)
        ops.gemm(POOLb, st.shadow(m.text_projection.weight), epilogue=ops.EPI_F32, out=pooled)
        save.B, save.S, save.ids, save.causal = B, S, ids, mask_u8 is None
        return (pooled, XF), save

    def backward(self, save, dpooled, dXF):
        m, s, st = self.mod, self.s, self.store
        d, B, S = s.d, save.B, save.S
        M = B * S
        bf, f32 = torch.bfloat16, torch.float32
        emb = m.embeddings
        G, Gb, _ = s.start_backward(save, M, None, None, _f32(dXF, (M, d)))
        if dpooled is not None:
            ln = getattr(m, "ln_final", None)
            W = m.text_projection.weight
            dPb = ops.cast_bf16(_f32(dpooled, (B, W.shape[0])))
            POOLb = save.get("ctxt.POOLb", (B, d), bf)
            ops.gemm(dPb, POOLb, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(W), accumulate=True)
            dL = torch.empty((B, d), device=s.device, dtype=f32)
            ops.gemm(dPb, st.shadow(W), b_mn=True, epilogue=ops.EPI_F32, out=dL)
            if ln is not None:
                ops.layernorm_bwd(save.get("ctxt.XSEL", (B, d), f32), None, dL, save.get("ctxt.mL", (B,), f32),
                                  save.get("ctxt.rL", (B,), f32), ln.weight, G, G, None, st.grad(ln.weight),
                                  st.grad(ln.bias), B, d, row_idx=self._idx, rows_per_group=S)
            else:
                ops.scatter_rows_add(dL, G, B, S, S - 1, d)
            ops.cast_bf16(G, Gb)
        s.stack.causal = save.causal
        G = s.stack.backward(G, Gb, B, S, top_bias_done=False, save=save)
        ops.batch_sum(G, st.grad(emb.position_embeddings), B, S * d, S * d)
        ops.batch_sum(G.view(-1)[(S - 1) * d:], st.grad(emb.cls_embedding), B, S * d, d)
        tok = emb.token_embeddings
        Gtok = G.view(B, S, d)[:, :S - 1].contiguous().view(-1, d)       # rows of the real tokens (slice copy: plumbing)
        ops.scatter_rows_idx_add(Gtok, save.ids.reshape(-1).contiguous(), st.grad(tok.weight), d)
        if tok.padding_idx is not None:
            ops.zero_(st.grad(tok.weight)[tok.padding_idx])
        return ()


class MultimodalDecoderTrainRuntime:
    """CoCaMultimodalDecoder up to its final LayerNorm (models/coca/multimodal_decoder.py:86-108); the vocabulary
    projection + cross-entropy is ``LinearCrossEntropyFunction``.  Inputs: texts [B, S, d], images [B, Si, d_kv]."""

    def __init__(self, mod: nn.Module):
        self.mod = mod
        self.s = _Stack(mod, mod.transformer_decoder.layer, "cmm", causal=True)
        self.store = self.s.store

    def forward(self, data, diff):
        texts, images = diff
        s, st = self.s, self.store
        d = s.d
        B, S, _ = texts.shape
        _, Si, dv = images.shape
        st.refresh()
        save = Workspace(s.device)
        X0 = torch.empty((B * S, d), device=s.device, dtype=torch.float32)
        X0.view(B, S, d).copy_(texts)
        enc = save.get("cmm.ENC", (B * Si, dv), torch.bfloat16)
        ops.cast_bf16(images.contiguous().float().view(-1), enc.view(-1))
        XM, Y = s.stack.forward(X0, B, S, True, save=save, enc=enc, S_enc=Si)
        XF, LAST = s.finish(XM, Y, B * S, self.mod.transformer_decoder.final_layer_norm, save)
        save.B, save.S, save.Si, save.dv = B, S, Si, dv
        return ((LAST if LAST is not None else XF),), save

    def backward(self, save, dOUT):
        s = self.s
        d, B, S = s.d, save.B, save.S
        M = B * S
        fln = self.mod.transformer_decoder.final_layer_norm
        dOUT = _f32(dOUT, (M, d))
        G, Gb, done = s.start_backward(save, M, fln, dOUT if fln is not None else None, None if fln is not None else dOUT)
        G = s.stack.backward(G, Gb, B, S, top_bias_done=done, save=save)
        return (G.view(B, S, d).clone(), save.dENC.view(B, save.Si, save.dv))


# ---------------------------------------------------------------------------------------------------------------------
class RuntimeFunction(torch.autograd.Function):
    """One training forward of a runtime above.  inputs: (runtime, data, n_diff, *diff_inputs, *parameters)."""

    @staticmethod
    def forward(ctx, rt, data, n_diff, *tensors):
        ctx.set_materialize_grads(False)
        outs, save = rt.forward(data, tensors[:n_diff])
        ctx.rt, ctx.save, ctx.n_diff = rt, save, n_diff
        ctx.need = ctx.needs_input_grad[3 + n_diff:]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        rt, save = ctx.rt, ctx.save
        if save is None:
            raise MMBError("this forward was already back-propagated (its activations are freed)")
        st = rt.store
        st.zero_grads()
        in_grads = rt.backward(save, *douts)
        ctx.save = None
        g = st.g.clone()
        grads = []
        for p, need in zip(st.params, ctx.need):
            o = st.off[id(p)]
            grads.append(g[o:o + p.numel()].view(p.shape) if need else None)
        return (None, None, None, *in_grads, *grads)


def run(rt, data, diff: Sequence[torch.Tensor] = ()) -> Tuple[torch.Tensor, ...]:
    return RuntimeFunction.apply(rt, data, len(diff), *diff, *rt.store.params)


class LinearF32Function(torch.autograd.Function):
    """y = x @ W^T (+ b) for a 2-D fp32 x, tensor-core GEMMs both ways.  The output width is padded to a multiple of 8
    internally (bf16 rows of d y must be 16-byte multiples for the TMA operands of the backward GEMMs)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        N, K = weight.shape
        Np = (N + 7) // 8 * 8
        dev = x.device
        xb = ops.cast_bf16(x.detach().contiguous().float())
        wb = ops.cast_bf16(weight.detach().contiguous())
        bp = bias.detach().float().contiguous() if bias is not None else None
        if Np != N:
            wp = torch.zeros((Np, K), device=dev, dtype=torch.bfloat16)
            wp[:N].copy_(wb)
            wb = wp
            if bp is not None:
                b2 = torch.zeros(Np, device=dev, dtype=torch.float32)
                b2[:N].copy_(bp)
                bp = b2
        out = torch.empty((xb.shape[0], Np), device=dev, dtype=torch.float32)
        ops.gemm(xb, wb, bias=bp, epilogue=ops.EPI_F32, out=out)
        ctx.save_for_backward(xb, wb)
        ctx.has_bias, ctx.N = bias is not None, N
        return out[:, :N]

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        N, Np = ctx.N, wb.shape[0]
        dyf = dy.contiguous().float()
        if Np != N:
            pad = torch.zeros((dyf.shape[0], Np), device=dy.device, dtype=torch.float32)
            pad[:, :N].copy_(dyf)
            dyf = pad
        dyb = ops.cast_bf16(dyf)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(xb.shape, device=dy.device, dtype=torch.float32)
            ops.gemm(dyb, wb, b_mn=True, epilogue=ops.EPI_F32, out=dx)
        if ctx.needs_input_grad[1]:
            gw = torch.empty(wb.shape, device=dy.device, dtype=torch.float32)
            ops.gemm(dyb, xb, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=gw)
            dW = gw[:N]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = torch.zeros(Np, device=dy.device, dtype=torch.float32)
            ops.colsum_bf16(dyb, gb, dyb.shape[0], Np, Np)
            db = gb[:N]
        return dx, dW, db


class LinearCrossEntropyFunction(torch.autograd.Function):
    """mean CrossEntropy(ignore_index)(hidden @ W^T (+ b), labels) — the captioning head (coca_model.py:443-454).
    Training materialises the fp32 logits [M, V] once (saved for the backward); the label-free forward path of
    CoCaForPretraining keeps the fused statistics GEMM."""

    @staticmethod
    def forward(ctx, hidden, weight, bias, labels, ignore_index):
        dev = hidden.device
        hb = ops.cast_bf16(hidden.detach().contiguous().float())
        M, d = hb.shape
        V = weight.shape[0]
        Vp = (V + 7) // 8 * 8
        wb = ops.cast_bf16(weight.detach().contiguous())
        if Vp != V:
            wp = torch.zeros((Vp, d), device=dev, dtype=torch.bfloat16)
            wp[:V].copy_(wb)
            wb = wp
        bp = None
        if bias is not None:
            bp = torch.zeros(Vp, device=dev, dtype=torch.float32)
            bp[:V].copy_(bias.detach().float())
        logits = torch.empty((M, Vp), device=dev, dtype=torch.float32)
        ops.gemm(hb, wb, bias=bp, epilogue=ops.EPI_F32, out=logits)
        lab = labels.reshape(-1).contiguous().long()
        accum = torch.zeros(2, device=dev, dtype=torch.float32)
        ops.ce_labels(logits[:, :V], lab, 1, ignore_index, M, V, None, accum)
        ctx.save_for_backward(hb, wb, logits, lab, accum)
        ctx.meta = (V, Vp, int(ignore_index), bias is not None)
        return accum[0] / accum[1]

    @staticmethod
    def backward(ctx, dloss):
        hb, wb, logits, lab, accum = ctx.saved_tensors
        V, Vp, ignore_index, has_bias = ctx.meta
        M, d = hb.shape
        dev = hb.device
        dlog = torch.zeros((M, Vp), device=dev, dtype=torch.bfloat16)
        ops.ce_labels_bwd(logits[:, :V], lab, 1, ignore_index, M, V, accum, 1.0, dlog[:, :V],
                          gscale=dloss.detach().float().reshape(1).contiguous())
        dh = dW = db = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty((M, d), device=dev, dtype=torch.float32)
            ops.gemm(dlog, wb, b_mn=True, epilogue=ops.EPI_F32, out=dh)
        if ctx.needs_input_grad[1]:
            gw = torch.empty((Vp, d), device=dev, dtype=torch.float32)
            ops.gemm(dlog, hb, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=gw, splits=ops.wgrad_splits(Vp, d, M))
            dW = gw[:V]
        if has_bias and ctx.needs_input_grad[2]:
            gb = torch.zeros(Vp, device=dev, dtype=torch.float32)
            ops.colsum_bf16(dlog, gb, M, Vp, Vp)
            db = gb[:V]
        return dh, dW, db, None, None


def linear_f32(x: torch.Tensor, linear: nn.Linear) -> torch.Tensor:
    """linear(x) over the last dim of x (any leading shape)."""
    lead = x.shape[:-1]
    y = LinearF32Function.apply(x.reshape(-1, x.shape[-1]), linear.weight, linear.bias)
    return y.view(*lead, -1)


def linear_cross_entropy(hidden: torch.Tensor, linear: nn.Linear, labels: torch.Tensor, ignore_index: int) -> torch.Tensor:
    return LinearCrossEntropyFunction.apply(hidden.reshape(-1, hidden.shape[-1]), linear.weight, linear.bias, labels,
                                            ignore_index)


class LayersTrainRuntime:
    """A standalone pre-norm `TransformerEncoderLayer` or `TransformerEncoder` (modules/layers/transformer.py:31-259) called
    on its own under autograd: hidden_states [B, S, d] in (differentiable), residual stream (and final LayerNorm) out."""

    def __init__(self, owner: nn.Module, layers, final_ln: Optional[nn.Module]):
        self.s = _Stack(owner, layers, "lyr", causal=False)
        self.store = self.s.store
        self.final_ln = final_ln

    def forward(self, data, diff):
        (mask_u8,) = data
        (x,) = diff
        s = self.s
        B, S, d = x.shape
        self.store.refresh()
        save = Workspace(s.device)
        X0 = torch.empty((B * S, d), device=s.device, dtype=torch.float32)
        X0.view(B, S, d).copy_(x)
        XM, Y = s.stack.forward(X0, B, S, True, save=save, mask3=mask_u8)
        XF, LAST = s.finish(XM, Y, B * S, self.final_ln, save)
        save.B, save.S = B, S
        self.last_hidden = ([X0.view(B, S, d)] + [save.bufs[f"lyr.XA.{l}"].view(B, S, d) for l in range(1, s.L)]
                            + [XF.view(B, S, d)])
        return ((LAST if LAST is not None else XF),), save

    def backward(self, save, dOUT):
        s = self.s
        d, B, S = s.d, save.B, save.S
        M = B * S
        fln = self.final_ln
        dOUT = _f32(dOUT, (M, d))
        G, Gb, done = s.start_backward(save, M, fln, dOUT if fln is not None else None, None if fln is not None else dOUT)
        G = s.stack.backward(G, Gb, B, S, top_bias_done=done, save=save)
        return (G.view(B, S, d).clone(),)


def standalone_layers(owner: nn.Module, layers, final_ln, hidden_states: torch.Tensor, mask_u8):
    """-> (output [B, S, d] with autograd history, hidden_states list) for a standalone layer / encoder call."""
    ids = [(id(p), p.device) for p in owner.parameters()]
    rt = getattr(owner, "_mmb_trt", None)
    if rt is None or getattr(owner, "_mmb_trt_ids", None) != ids:
        rt = LayersTrainRuntime(owner, layers, final_ln)
        object.__setattr__(owner, "_mmb_trt", rt)
        object.__setattr__(owner, "_mmb_trt_ids", ids)
    (out,) = run(rt, (mask_u8,), (hidden_states.float(),))
    hidden, rt.last_hidden = rt.last_hidden, None
    return out.view(hidden_states.shape), hidden
