"""Training runtime of the FLAVA encoders (BASELINE.json config 3 as a training step): forward that keeps what the
backward needs + the explicit backward schedule, behind torch.autograd Functions so that the drop-in modules train
with ``loss.backward()`` exactly like the reference's (models/flava/model.py:127-298 under autograd).

The layer stack is the CLIP towers' ``engine.TransformerStack`` (same kernels, same fused schedule): the separate
query / key / value Linears are presented to it as one packed in-projection (``ParamStore.pack``), the MLP activation
is the exact-erf GELU epilogue pair, and the text tower's key-padding mask goes to the masked attention kernels
(forward and fused single-pass backward).  What is specific to FLAVA is on either side of the stack:

  image  : im2col + patch GEMM (+bias) -> [cls | mask_token or patch] + pos      bwd: mmb_vit_assemble_bwd, batch sums,
           patch-projection weight / bias gradients                                   (image_encoder.py:139-175)
  text   : LayerNorm(word + pos + type) with pad-derived key mask                 bwd: mmb_bert_embed_ln_bwd (recomputes
           the pre-norm sum, scatter-adds into the three tables)                      (text_embedding.py:70-104)
  mm     : two projections -> [cls | image | text]                              bwd: mmb_split_tokens_cast, projection
           gradients, gradients w.r.t. both incoming hidden states                    (model.py:283-298)
  all    : final LayerNorm -> last_hidden_state; its input is hidden_states[-1], which the multimodal encoder consumes,
           so BOTH are differentiable outputs.  Pooler / `linear(last_hidden_state[:, 0])` projections are
           ``FirstTokenLinearFunction`` (gather, GEMM, tanh; backward scatters into the dense gradient).

One encoder instance runs twice per pre-training step (unmasked + masked inputs): every training forward keeps its
activations in its OWN Workspace (held by the autograd node, freed after its backward), so any number of forwards may
be in flight.  ``hidden_states[1:-1]`` of a training forward are views of those saved buffers: values are the
reference's, but they carry no autograd history (nothing in the library differentiates through them).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops
from ._lib import MMBError
from .engine import ParamStore, TransformerStack, Workspace
from .modules.layers.transformer import TransformerOutput


def wants_grad(*mods: Optional[nn.Module]) -> bool:
    """True when the caller expects an autograd graph: grad mode on and some parameter of `mods` trainable."""
    if not torch.is_grad_enabled():
        return False
    return any(p.requires_grad for m in mods if m is not None for p in m.parameters())


class FlavaTrainStack:
    """ParamStore (packed q/k/v order) + TransformerStack + final LayerNorm of one FLAVA encoder."""

    def __init__(self, owner: nn.Module, encoder: nn.Module, layernorm: nn.Module, prefix: str,
                 extra: Sequence[nn.Module] = ()):
        layers = list(encoder.layer)
        l0 = layers[0]
        if not l0.norm_first:
            raise MMBError("only pre-norm (norm_first=True) FLAVA layers are on the accelerated path")
        if not isinstance(l0.feedforward.model[1], nn.GELU):
            raise MMBError("unsupported MLP activation (FLAVA uses nn.GELU)")
        d, H = l0.attention.dim_q, l0.attention.n_head
        ff = l0.feedforward.model[0].weight.shape[0]
        params: List[nn.Parameter] = []
        for layer in layers:   # q / k / v weights, then biases, consecutive -> packable
            at = layer.attention
            params += [at.query.weight, at.key.weight, at.value.weight, at.query.bias, at.key.bias, at.value.bias]
        seen = {id(p) for p in params}
        for m in (owner, *extra):
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        self.store = ParamStore(params)
        self.device = self.store.device
        st = self.store
        adapters = []
        for layer in layers:   # the attribute names TransformerStack reads (torch.nn.TransformerEncoderLayer layout)
            at, mlp = layer.attention, layer.feedforward.model
            attn = SimpleNamespace(in_proj_weight=st.pack([at.query.weight, at.key.weight, at.value.weight]),
                                   in_proj_bias=st.pack([at.query.bias, at.key.bias, at.value.bias], fp32=True),
                                   out_proj=at.output, num_heads=H)
            adapters.append(SimpleNamespace(self_attn=attn, norm1=layer.attention_layernorm,
                                            norm2=layer.feedforward_layernorm, linear1=mlp[0], linear2=mlp[-1]))
        self.ws = Workspace(self.device)   # scratch shared by all calls (stream-ordered)
        self.stack = TransformerStack(adapters, st, self.ws, d=d, heads=H, ff=ff, causal=False, act=ops.ACT_GELU_ERF,
                                      prefix=prefix)
        self.layernorm, self.prefix = layernorm, prefix
        self.d, self.H, self.L = d, H, len(layers)

    def forward(self, X0: torch.Tensor, B: int, S: int, kmask: Optional[torch.Tensor], save: Workspace):
        """Returns (LAST, XF, hidden_states): fp32 [B*S, d] each; LAST = layernorm(XF), XF = hidden_states[-1]."""
        d, ln, pfx = self.d, self.layernorm, self.prefix
        M = B * S
        f32 = torch.float32
        if kmask is not None and S > 256:
            raise MMBError("training with a key-padding mask is implemented for sequence lengths <= 256")
        XM, Y = self.stack.forward(X0, B, S, True, kmask=kmask, save=save)
        XF = torch.empty((M, d), device=self.device, dtype=f32)
        LAST = torch.empty((M, d), device=self.device, dtype=f32)
        ops.add_layernorm_fwd(XM, Y, XF, None, LAST, ln.weight, ln.bias, save.get(f"{pfx}.mF", (M,), f32),
                              save.get(f"{pfx}.rF", (M,), f32), M, d, ln.eps)
        save.XF, save.B, save.S = XF, B, S
        hidden = [X0.view(B, S, d)]
        hidden += [save.bufs[f"{pfx}.XA.{l}"].view(B, S, d) for l in range(1, self.L)]
        hidden.append(XF.view(B, S, d))
        return LAST, XF, hidden

    def backward(self, save: Workspace, dLAST: Optional[torch.Tensor], dXF: Optional[torch.Tensor]) -> torch.Tensor:
        """Gradient w.r.t. X0 (fp32 [B*S, d], scratch: consume before the next backward of this encoder); parameter
        gradients are accumulated into the store's flat buffer."""
        d, ln, pfx, st = self.d, self.layernorm, self.prefix, self.store
        B, S = save.B, save.S
        M = B * S
        f32, bf = torch.float32, torch.bfloat16
        G = self.ws.get(f"{pfx}.G", (M, d), f32)
        Gb = self.ws.get(f"{pfx}.Gb", (M, d), bf)
        if dLAST is None:   # only hidden_states[-1] was used downstream: LayerNorm backward of a zero gradient
            dLAST = torch.zeros((M, d), device=self.device, dtype=f32)
        ops.layernorm_bwd(save.XF, None, dLAST, save.get(f"{pfx}.mF", (M,), f32), save.get(f"{pfx}.rF", (M,), f32),
                          ln.weight, dXF, G, Gb, st.grad(ln.weight), st.grad(ln.bias), M, d,
                          gsum=self.stack.top_bias_grad())
        return self.stack.backward(G, Gb, B, S, top_bias_done=True, save=save)


def _f32c(t: Optional[torch.Tensor], shape) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.contiguous().float().view(shape)


class FlavaImageTrainRuntime:
    def __init__(self, mod: nn.Module):
        self.mod = mod
        self.ts = FlavaTrainStack(mod, mod.encoder, mod.layernorm, "fimg")
        self.store = self.ts.store

    def diff_inputs(self, data) -> Tuple[torch.Tensor, ...]:
        return ()

    def forward(self, data, diff):
        pixel_values, image_patches_mask = data
        emb, ts, st = self.mod.embeddings, self.ts, self.store
        d = ts.d
        conv = emb.patch_embeddings.projection
        ps = conv.weight.shape[2]
        image = pixel_values.contiguous().float()
        B, _, Hh, Ww = image.shape
        P = (Hh // ps) * (Ww // ps)
        S = P + 1
        K = 3 * ps * ps
        Kp = -(-K // 8) * 8
        bf, f32 = torch.bfloat16, torch.float32
        st.refresh()
        save = Workspace(ts.device)
        PATCH = save.get("fimg.PATCH", (B * P, Kp), bf)[:, :K]
        PO = ts.ws.get("fimg.PO", (B * P, d), bf)
        X0 = torch.empty((B * S, d), device=image.device, dtype=f32)
        ops.im2col(image, ps, PATCH)
        w = st.shadow2d(conv.weight)
        if Kp != K:
            wp = ts.ws.get("fimg.WCONV", (d, Kp), bf)[:, :K]
            wp.copy_(w)
            w = wp
        ops.gemm(PATCH, w, bias=conv.bias, out=PO)
        pm = None
        if image_patches_mask is not None and emb.mask_token is not None:
            pm = image_patches_mask.reshape(B, P).to(torch.uint8).contiguous()
        ops.vit_assemble_fwd(PO, emb.cls_token, emb.position_embeddings, emb.mask_token if pm is not None else None, pm, X0,
                             B, S, d)
        save.pm, save.P, save.K = pm, P, K
        LAST, XF, hidden = ts.forward(X0, B, S, None, save)
        return LAST, XF, hidden, save

    def backward(self, save, dLAST, dXF):
        emb, ts, st = self.mod.embeddings, self.ts, self.store
        d, B, S, P, K = ts.d, save.B, save.S, save.P, save.K
        conv = emb.patch_embeddings.projection
        G = ts.backward(save, dLAST, dXF)
        ops.batch_sum(G, st.grad(emb.position_embeddings), B, S * d, S * d)
        ops.batch_sum(G, st.grad(emb.cls_token), B, S * d, d)
        DP = ts.ws.get("fimg.DP", (B * P, d), torch.bfloat16)
        ops.vit_assemble_bwd(G, save.pm, DP, st.grad(emb.mask_token) if save.pm is not None else None, B, S, d, True)
        PATCH = save.get("fimg.PATCH", (B * P, -(-K // 8) * 8), torch.bfloat16)[:, :K]
        ops.gemm(DP, PATCH, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad2d(conv.weight),
                 splits=ops.wgrad_splits(d, PATCH.shape[1], B * P), accumulate=True)
        ops.colsum_bf16(DP, st.grad(conv.bias), B * P, d, d)
        return ()


class FlavaTextTrainRuntime:
    def __init__(self, mod: nn.Module):
        self.mod = mod
        self.ts = FlavaTrainStack(mod, mod.encoder, mod.layernorm, "ftxt")
        self.store = self.ts.store

    def diff_inputs(self, data):
        return ()

    def forward(self, data, diff):
        input_ids, attention_mask, token_type_ids = data
        emb, ts, st = self.mod.embeddings, self.ts, self.store
        d = ts.d
        ids = input_ids.long().contiguous()
        B, S = ids.shape
        if S > emb.position_embeddings.weight.shape[0]:
            raise ValueError(f"sequence length {S} exceeds max_position_embeddings")
        st.refresh()
        save = Workspace(ts.device)
        X0 = torch.empty((B * S, d), device=ids.device, dtype=torch.float32)
        KM = save.get("ftxt.KM", (B * S,), torch.uint8)
        tt = token_type_ids.long().contiguous() if token_type_ids is not None else None
        V = emb.word_embeddings.weight.shape[0]
        ops.bert_embed_ln_fwd(ids, tt, emb.word_embeddings.weight, emb.position_embeddings.weight,
                              emb.token_type_embeddings.weight, emb.layer_norm.weight, emb.layer_norm.bias, X0, KM,
                              emb.pad_token_id, B, S, d, V, emb.layer_norm.eps)
        if attention_mask is not None:
            if attention_mask.dim() != 2:
                raise NotImplementedError("only [batch, seq_len] padding masks are supported on the accelerated path")
            KM = (attention_mask != 0).to(torch.uint8).contiguous().view(-1)
        save.ids, save.tt, save.V = ids, tt, V
        LAST, XF, hidden = ts.forward(X0, B, S, KM, save)
        return LAST, XF, hidden, save

    def backward(self, save, dLAST, dXF):
        emb, ts, st = self.mod.embeddings, self.ts, self.store
        d, B, S = ts.d, save.B, save.S
        G = ts.backward(save, dLAST, dXF)
        word = emb.word_embeddings
        ops.bert_embed_ln_bwd(save.ids, save.tt, word.weight, emb.position_embeddings.weight,
                              emb.token_type_embeddings.weight, emb.layer_norm.weight, G, st.grad(word.weight),
                              st.grad(emb.position_embeddings.weight), st.grad(emb.token_type_embeddings.weight),
                              st.grad(emb.layer_norm.weight), st.grad(emb.layer_norm.bias), B, S, d, save.V,
                              emb.layer_norm.eps)
        if word.padding_idx is not None:   # nn.Embedding(padding_idx): that row receives no gradient
            ops.zero_(st.grad(word.weight)[word.padding_idx])
        return ()


class FlavaMMTrainRuntime:
    """[cls | image_to_mm(image_hidden) | text_to_mm(text_hidden)] -> stack.  The two projection Linears belong to
    FLAVAModel, not to the multimodal encoder; they live in this runtime's ParamStore (image_proj / text_proj None:
    the module was called directly with an already fused token sequence)."""

    def __init__(self, mod: nn.Module, image_proj: Optional[nn.Linear], text_proj: Optional[nn.Linear]):
        self.mod, self.image_proj, self.text_proj = mod, image_proj, text_proj
        extra = [m for m in (image_proj, text_proj) if m is not None]
        self.ts = FlavaTrainStack(mod, mod.encoder, mod.layernorm, "fmm", extra=extra)
        self.store = self.ts.store

    def forward(self, data, diff):
        ts, st = self.ts, self.store
        d = ts.d
        bf, f32 = torch.bfloat16, torch.float32
        cls = self.mod.cls_token
        off = 1 if cls is not None else 0
        st.refresh()
        save = Workspace(ts.device)
        if self.image_proj is None:   # direct call: hidden_states [B, S, d] already fused
            (hs,) = diff
            B, Sa, dd = hs.shape
            hs = hs.contiguous().float()
            S = Sa + off
            X0 = torch.empty((B * S, d), device=hs.device, dtype=f32)
            ops.concat_tokens(cls, hs, hs, X0, B, Sa, 0, d)
            save.Si, save.St = Sa, 0
        else:
            image_hidden, text_hidden = diff
            B, Si, di = image_hidden.shape
            Bt, St, dt = text_hidden.shape
            if B != Bt:
                raise ValueError(f"batch mismatch between image ({B}) and text ({Bt}) hidden states")
            Ib = save.get("fmm.Ib", (B * Si, di), bf)
            Tb = save.get("fmm.Tb", (B * St, dt), bf)
            ops.cast_bf16(image_hidden.contiguous().float().view(-1), Ib.view(-1))
            ops.cast_bf16(text_hidden.contiguous().float().view(-1), Tb.view(-1))
            Pi = ts.ws.get("fmm.Pi", (B * Si, d), f32)
            Pt = ts.ws.get("fmm.Pt", (B * St, d), f32)
            ops.gemm(Ib, st.shadow(self.image_proj.weight), bias=self.image_proj.bias, epilogue=ops.EPI_F32, out=Pi)
            ops.gemm(Tb, st.shadow(self.text_proj.weight), bias=self.text_proj.bias, epilogue=ops.EPI_F32, out=Pt)
            S = Si + St + off
            X0 = torch.empty((B * S, d), device=image_hidden.device, dtype=f32)
            ops.concat_tokens(cls, Pi, Pt, X0, B, Si, St, d)
            save.Si, save.St, save.di, save.dt = Si, St, di, dt
        LAST, XF, hidden = ts.forward(X0, B, S, None, save)
        return LAST, XF, hidden, save

    def backward(self, save, dLAST, dXF):
        ts, st = self.ts, self.store
        d, B, S, Si, St = ts.d, save.B, save.S, save.Si, save.St
        bf, f32 = torch.bfloat16, torch.float32
        cls = self.mod.cls_token
        has_cls = cls is not None
        G = ts.backward(save, dLAST, dXF)
        if has_cls:
            ops.batch_sum(G, st.grad(cls), B, S * d, d)
        if self.image_proj is None:
            dH = G.view(B, S, d)[:, (1 if has_cls else 0):].clone()   # strided slice copy: plumbing
            return (dH,)
        dPi = ts.ws.get("fmm.dPi", (B * Si, d), bf)
        dPt = ts.ws.get("fmm.dPt", (B * St, d), bf)
        ops.split_tokens_cast(G, dPi, dPt, B, Si, St, d, has_cls)
        outs = []
        for dP, lin, key, n, din in ((dPi, self.image_proj, "fmm.Ib", B * Si, save.di),
                                     (dPt, self.text_proj, "fmm.Tb", B * St, save.dt)):
            Xb = save.get(key, (n, din), bf)
            ops.gemm(dP, Xb, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(lin.weight),
                     splits=ops.wgrad_splits(d, din, n), accumulate=True)
            ops.colsum_bf16(dP, st.grad(lin.bias), n, d, d)
            dX = torch.empty((n, din), device=ts.device, dtype=f32)
            ops.gemm(dP, st.shadow(lin.weight), b_mn=True, epilogue=ops.EPI_F32, out=dX)
            outs.append(dX)
        return (outs[0].view(B, Si, save.di), outs[1].view(B, St, save.dt))


class FlavaEncodeFunction(torch.autograd.Function):
    """One training forward of a FLAVA encoder.  inputs: (runtime, data, n_diff, *diff_inputs, *parameters);
    outputs: (last_hidden_state, hidden_states[-1]) as fp32 [B*S, d]."""

    @staticmethod
    def forward(ctx, rt, data, n_diff, *tensors):
        diff = tensors[:n_diff]
        ctx.set_materialize_grads(False)   # an unused output arrives as None in backward, not as a zero tensor
        LAST, XF, hidden, save = rt.forward(data, diff)
        rt.last_hidden = hidden
        ctx.rt, ctx.save, ctx.n_diff, ctx.n_par = rt, save, n_diff, len(tensors) - n_diff
        ctx.need = ctx.needs_input_grad[3 + n_diff:]
        return LAST, XF

    @staticmethod
    def backward(ctx, dLAST, dXF):
        rt, save = ctx.rt, ctx.save
        if save is None:
            raise MMBError("this encoder forward was already back-propagated (its activations are freed)")
        st = rt.store
        M = save.B * save.S
        st.zero_grads()
        in_grads = rt.backward(save, _f32c(dLAST, (M, -1)), _f32c(dXF, (M, -1)))
        ctx.save = None
        g = st.g.clone()
        grads = []
        for p, need in zip(st.params, ctx.need):
            o = st.off[id(p)]
            grads.append(g[o:o + p.numel()].view(p.shape) if need else None)
        return (None, None, None, *in_grads, *grads)


def run_encoder(rt, data, diff: Sequence[torch.Tensor] = ()):
    """-> (LAST [M,d], XF [M,d], hidden_states list) with LAST / XF attached to the autograd graph."""
    LAST, XF = FlavaEncodeFunction.apply(rt, data, len(diff), *diff, *rt.store.params)
    hidden = rt.last_hidden
    rt.last_hidden = None
    return LAST, XF, hidden


class FirstTokenLinearFunction(torch.autograd.Function):
    """y = [tanh](linear(x[:, 0, :]))  (Pooler: modules/losses/flava.py:84-97; cls projections: model.py:244-246,
    261-263).  x fp32 [B, S, d].  Backward: weight / bias gradients by GEMM / column sum, dx scattered into row 0."""

    @staticmethod
    def forward(ctx, x, weight, bias, use_tanh):
        B, S, d = x.shape
        E = weight.shape[0]
        xf = x.contiguous().float()
        CLSb = torch.empty((B, d), device=x.device, dtype=torch.bfloat16)
        ops.gather_rows_cast(xf.view(B * S, d), CLSb, B, S, 0, d)
        wb = ops.cast_bf16(weight.detach().contiguous())
        out = torch.empty((B, E), device=x.device, dtype=torch.float32)
        ops.gemm(CLSb, wb, bias=bias.detach() if bias is not None else None, epilogue=ops.EPI_F32, out=out)
        if use_tanh:
            ops.tanh_(out)
        ctx.save_for_backward(CLSb, wb, out if use_tanh else None)
        ctx.shape, ctx.use_tanh, ctx.has_bias = (B, S, d, E), use_tanh, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        CLSb, wb, y = ctx.saved_tensors
        B, S, d, E = ctx.shape
        dev = dout.device
        dof = dout.contiguous().float()
        if ctx.use_tanh:
            dpre = torch.empty((B, E), device=dev, dtype=torch.bfloat16)
            ops.tanh_bwd(dof, y, None, dpre)
        else:
            dpre = ops.cast_bf16(dof)
        dW = db = dx = None
        if ctx.needs_input_grad[1]:
            dW = torch.empty((E, d), device=dev, dtype=torch.float32)
            ops.gemm(dpre, CLSb, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=dW)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(E, device=dev, dtype=torch.float32)
            ops.colsum_bf16(dpre, db, B, E, E)
        if ctx.needs_input_grad[0]:
            dcls = torch.empty((B, d), device=dev, dtype=torch.float32)
            ops.gemm(dpre, wb, b_mn=True, epilogue=ops.EPI_F32, out=dcls)
            dx = torch.zeros((B, S, d), device=dev, dtype=torch.float32)
            ops.scatter_rows_add(dcls, dx.view(B * S, d), B, S, 0, d)
        return dx, dW, db, None


def first_token_linear(x: torch.Tensor, linear: nn.Linear, use_tanh: bool = False) -> torch.Tensor:
    return FirstTokenLinearFunction.apply(x, linear.weight, linear.bias, use_tanh)


def encoder_output(rt, data, diff, pooler: Optional[nn.Module]) -> TransformerOutput:
    """Training-mode TransformerOutput of one encoder call (pooler applied through FirstTokenLinearFunction)."""
    LAST, XF, hidden = run_encoder(rt, data, diff)
    B, S, d = hidden[0].shape
    last = LAST.view(B, S, d)
    hidden = list(hidden[:-1]) + [XF.view(B, S, d)]   # hidden_states[-1] is the differentiable output
    pooled = first_token_linear(last, pooler.dense, use_tanh=True) if pooler is not None else None
    return TransformerOutput(last_hidden_state=last, pooler_output=pooled, hidden_states=hidden, attentions=None)
