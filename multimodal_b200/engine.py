"""Host-side runtime of the dual-encoder path: parameter shadows, activation workspaces and the explicit
forward / backward schedules that drive the sm_100a kernels (multimodal_b200.ops).

Nothing here computes: every tensor op is a C-ABI kernel launch on the current CUDA stream.  The schedules follow
the reference call stack:
  torch/nn/modules/transformer.py:946-951 (pre-norm layer), torch/nn/functional.py:6478-6690 (MHA),
  models/clip/image_encoder.py:82-113, models/clip/text_encoder.py:113-134.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import ops
from ._lib import MMBError

_ALIGN = 64  # elements; keeps every shadow / grad slice 128 B aligned (TMA needs 16 B)

# bf16 operand shadows are re-cast when a parameter's autograd version counter (or storage) changes.  In-place writes
# through `.data` (`w.data.copy_()`, EMA updates, `module.weight.data.normal_()`) do NOT bump that counter, so every
# shadow key also carries this process-wide epoch: `invalidate_weight_caches()` (exported from the package root, and
# called by the drop-in modules' load_state_dict / _apply hooks) forces a re-cast on the next forward.
_WEIGHT_EPOCH = [0]


def invalidate_weight_caches() -> None:
    _WEIGHT_EPOCH[0] += 1


def weight_epoch() -> int:
    return _WEIGHT_EPOCH[0]


def watch_module(mod: nn.Module) -> None:
    """Invalidate the shadows whenever `mod` (or a parent calling into it) reloads or re-homes its parameters."""
    if getattr(mod, "_mmb_watched", False):
        return
    mod._mmb_watched = True
    mod.register_load_state_dict_post_hook(lambda module, incompatible: invalidate_weight_caches())


def _require_cuda(dev: torch.device) -> None:
    if dev.type != "cuda":
        raise MMBError("multimodal_b200 modules must live on a CUDA device (no CPU path); call .cuda() first")


class ParamStore:
    """Flat bf16 shadow (tensor-core operand copies) + flat fp32 gradient buffer for a list of parameters."""

    def __init__(self, params: Sequence[nn.Parameter]):
        self.params: List[nn.Parameter] = list(params)
        if not self.params:
            raise MMBError("ParamStore: no parameters")
        dev = self.params[0].device
        _require_cuda(dev)
        self.device = dev
        self.off: Dict[int, int] = {}
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise MMBError("parameters must be fp32 (bf16 operand copies are made internally)")
            self.off[id(p)] = off
            off += -(-p.numel() // _ALIGN) * _ALIGN
        self.total = off
        self.wb = torch.empty(self.total, device=dev, dtype=torch.bfloat16)
        self.g = torch.zeros(self.total, device=dev, dtype=torch.float32)
        self._seen: Dict[int, tuple] = {}
        self.master: Optional[torch.Tensor] = None  # set by flatten_()
        self._shadow_fresh = False
        self._epoch = _WEIGHT_EPOCH[0]
        self._packs: List[list] = []   # [key tensor, parts, version seen]: fp32 concatenations kept current by refresh()
        self._keys: List[torch.Tensor] = []

    # -- views ------------------------------------------------------------------------------------------
    def shadow(self, p: nn.Parameter) -> torch.Tensor:
        o = self.off[id(p)]
        return self.wb[o:o + p.numel()].view(p.shape)

    def shadow2d(self, p: nn.Parameter) -> torch.Tensor:
        o = self.off[id(p)]
        return self.wb[o:o + p.numel()].view(p.shape[0], -1)

    def grad(self, p: nn.Parameter) -> torch.Tensor:
        o = self.off[id(p)]
        return self.g[o:o + p.numel()].view(p.shape)

    def grad2d(self, p: nn.Parameter) -> torch.Tensor:
        o = self.off[id(p)]
        return self.g[o:o + p.numel()].view(p.shape[0], -1)

    def pack(self, parts: Sequence[nn.Parameter], fp32: bool = False) -> torch.Tensor:
        """Present consecutive parameters as ONE tensor (rows concatenated): e.g. separate query / key / value Linears
        as the packed [3d, d] in-projection operand.  Returns a key tensor accepted by shadow() / grad() (views spanning
        all parts).  fp32=True: the key is a real fp32 concatenation, kept current by refresh(), usable as a kernel
        operand (biases); otherwise it is a shape-only placeholder."""
        offs = [self.off[id(p)] for p in parts]
        for a, b, p in zip(offs, offs[1:], parts):
            if b != a + p.numel():
                raise MMBError("ParamStore.pack: parts must be consecutive in the store with sizes that are multiples "
                               f"of {_ALIGN} elements")
        shape = (sum(p.shape[0] for p in parts),) + tuple(parts[0].shape[1:])
        if fp32:
            key = torch.empty(shape, device=self.device, dtype=torch.float32)
            self._packs.append([key, list(parts), None])
        else:
            key = torch.empty(shape, device="meta", dtype=torch.float32)
        self.off[id(key)] = offs[0]
        self._keys.append(key)   # keeps id(key) unique for the lifetime of the store
        return key

    def _refresh_packs(self, force: bool) -> None:
        for ent in self._packs:
            key, parts, seen = ent
            ver = tuple((p._version, p.data_ptr()) for p in parts) + (_WEIGHT_EPOCH[0],)
            if force or seen != ver:
                torch.cat([p.data.reshape(-1) for p in parts], out=key.view(-1))   # a few KB of biases: plumbing
                ent[2] = ver

    # -- maintenance ------------------------------------------------------------------------------------
    def flatten_(self) -> None:
        """Re-home every parameter into one flat fp32 master buffer (p.data and p.grad become views).  Enables the
        single-kernel fused optimizer / single NCCL all-reduce of multimodal_b200.train."""
        if self.master is not None:
            return
        self.master = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        with torch.no_grad():
            for p in self.params:
                o = self.off[id(p)]
                self.master[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.master[o:o + p.numel()].view(p.shape)
                p.grad = self.g[o:o + p.numel()].view(p.shape)
        self._shadow_fresh = False

    def refresh(self) -> None:
        """Make the bf16 shadows current (re-cast whatever changed since the last call)."""
        if self.master is not None:
            if not self._shadow_fresh or self._epoch != _WEIGHT_EPOCH[0]:
                ops.cast_bf16(self.master, self.wb)
                self._refresh_packs(True)
                self._shadow_fresh = True
                self._epoch = _WEIGHT_EPOCH[0]
            return
        for p in self.params:
            if p.device != self.device:
                raise MMBError("parameter moved to another device after the runtime was created")
            key = (p._version, p.data_ptr(), _WEIGHT_EPOCH[0])
            if self._seen.get(id(p)) != key:
                src = p.data if p.data.is_contiguous() else p.data.contiguous()
                ops.cast_bf16(src.view(-1), self.wb[self.off[id(p)]:self.off[id(p)] + p.numel()])
                self._seen[id(p)] = key
        self._refresh_packs(False)

    def mark_dirty(self) -> None:
        self._shadow_fresh = False

    def zero_grads(self) -> None:
        ops.zero_(self.g)


class Workspace:
    """Named, lazily allocated, reused device buffers."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}
        self.X0: Optional[torch.Tensor] = None      # set by TransformerStack.forward(training=True): layer-0 input
        self.kmask: Optional[torch.Tensor] = None   # ... and the key-padding mask / [B,S,S] mask / cross-attention
        self.mask3: Optional[torch.Tensor] = None   #     source that forward used
        self.enc: Optional[torch.Tensor] = None
        self.S_enc = 0
        self.dENC: Optional[torch.Tensor] = None    # set by backward: gradient w.r.t. the cross-attention source

    def get(self, name: str, shape, dtype) -> torch.Tensor:
        t = self.bufs.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or t.dtype != dtype or t.numel() < _numel(shape):
            t = torch.empty(_numel(shape), device=self.device, dtype=dtype)
            self.bufs[name] = t
        return t[:_numel(shape)].view(shape)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


class TransformerStack:
    """L pre-norm encoder layers (torch.nn.TransformerEncoderLayer parameter layout), QuickGELU or GELU MLP."""

    def __init__(self, layers: Sequence[nn.Module], store: ParamStore, ws: Workspace, *, d: int, heads: int, ff: int,
                 causal: bool, act: int, prefix: str):
        self.layers = list(layers)
        self.store, self.ws = store, ws
        self.d, self.H, self.ff, self.causal, self.act, self.prefix = d, heads, ff, causal, act, prefix
        if d % heads or d // heads != 64:
            raise MMBError(f"attention kernels support head_dim 64 only (got d={d}, heads={heads})")
        self.scale = 1.0 / 8.0
        self.L = len(self.layers)
        self.saved = False

    def _buf(self, name, l, shape, dtype, training):
        key = f"{self.prefix}.{name}.{l if training else 0}"
        return (self._save if training else self.ws).get(key, shape, dtype)

    def forward(self, X0: torch.Tensor, B: int, S: int, training: bool, kmask: Optional[torch.Tensor] = None,
                save: Optional["Workspace"] = None, mask3: Optional[torch.Tensor] = None,
                enc: Optional[torch.Tensor] = None, S_enc: int = 0):
        """X0: fp32 [B*S, d] residual stream entering layer 0.  Returns (XM_last fp32, Y bf16): the final residual
        stream is XM_last + Y (the add is fused into whichever LayerNorm consumes it).
        kmask: optional uint8 [B*S] key-padding mask (1 = attend).  mask3: optional uint8 [B, S, S] mask (general
        attention kernels).  enc / S_enc: bf16 [B*S_enc, d_kv] cross-attention source for layers that carry a
        `cross_attn` block (TransformerDecoderLayer: modules/layers/transformer.py:354-377).  save: the Workspace that receives the activations a
        training forward keeps for its backward (default: the stack's own — ONE in-flight training forward; callers
        that run the same stack several times before the backward pass a fresh Workspace per call)."""
        self._save = save if save is not None else self.ws
        self._save.kmask = kmask if training else None
        self._save.mask3 = mask3 if training else None
        self._save.enc, self._save.S_enc = (enc, S_enc) if training else (None, 0)
        if mask3 is not None and kmask is not None:
            raise MMBError("TransformerStack: pass either a key-padding mask or a [B, S, S] mask, not both")
        st, d, ff, H = self.store, self.d, self.ff, self.H
        M = B * S
        bf, f32 = torch.bfloat16, torch.float32
        Y = self.ws.get(f"{self.prefix}.Y", (M, d), bf)
        XA_prev, XM_prev = X0, None
        for l, layer in enumerate(self.layers):
            at = layer.self_attn
            LN1 = self._buf("LN1", l, (M, d), bf, training)
            QKV = self._buf("QKV", l, (M, 3 * d), bf, training)
            O = self._buf("O", l, (M, d), bf, training)
            LSE = self._buf("LSE", l, (B * H * S,), f32, training)
            XM = self._buf("XM", l, (M, d), f32, training)
            LN2 = self._buf("LN2", l, (M, d), bf, training)
            PRE = self._buf("PRE", l, (M, ff), bf, training)
            HACT = self._buf("HACT", l, (M, ff), bf, training)
            m1 = self._buf("m1", l, (M,), f32, training); r1 = self._buf("r1", l, (M,), f32, training)
            m2 = self._buf("m2", l, (M,), f32, training); r2 = self._buf("r2", l, (M,), f32, training)
            if l == 0:
                XA = X0
                ops.add_layernorm_fwd(XA, None, None, LN1, None, layer.norm1.weight, layer.norm1.bias, m1, r1, M, d,
                                      layer.norm1.eps)
            else:
                XA = self._buf("XA", l, (M, d), f32, training)
                ops.add_layernorm_fwd(XM_prev, Y, XA, LN1, None, layer.norm1.weight, layer.norm1.bias, m1, r1, M, d,
                                      layer.norm1.eps)
            ops.gemm(LN1, st.shadow(at.in_proj_weight), bias=at.in_proj_bias, out=QKV)
            if mask3 is not None:
                ops.attention_fwd_generic(QKV[:, :d], QKV[:, d:2 * d], QKV[:, 2 * d:], O, B=B, Sq=S, Skv=S, H=H, head_dim=64,
                                          bsq=S * 3 * d, bsk=S * 3 * d, bsv=S * 3 * d, bso=S * d, scale=self.scale,
                                          mask=mask3, mask_bs=S * S, mask_qs=S, causal=self.causal)
            elif kmask is not None:
                ops.attention_fwd_kmask(QKV, O, LSE, kmask, B, S, H, self.causal, self.scale)
            else:
                ops.attention_fwd(QKV, O, LSE, B, S, H, self.causal, self.scale)
            ops.gemm(O, st.shadow(at.out_proj.weight), bias=at.out_proj.bias, out=Y)
            ca = getattr(layer, "cross_attn", None)
            if ca is not None and enc is not None:
                # x1 = x + self-attention (XM);  x2 = x1 + cross-attention(LN_c(x1), enc) (XC);  the MLP reads LN2(x2)
                lnc = layer.norm_cross
                Se = S_enc
                LNC = self._buf("LNC", l, (M, d), bf, training)
                QC = self._buf("QC", l, (M, d), bf, training)
                KVC = self._buf("KVC", l, (B * Se, 2 * d), bf, training)
                OC = self._buf("OC", l, (M, d), bf, training)
                XC = self._buf("XC", l, (M, d), f32, training)
                mc = self._buf("mc", l, (M,), f32, training); rc = self._buf("rc", l, (M,), f32, training)
                ops.add_layernorm_fwd(XA, Y, XM, LNC, None, lnc.weight, lnc.bias, mc, rc, M, d, lnc.eps)
                ops.gemm(LNC, st.shadow(ca.q_w), bias=ca.q_b, out=QC)
                ops.gemm(enc, st.shadow(ca.kv_w), bias=ca.kv_b, out=KVC)
                ops.attention_fwd_generic(QC, KVC[:, :d], KVC[:, d:], OC, B=B, Sq=S, Skv=Se, H=H, head_dim=64, bsq=S * d,
                                          bsk=Se * 2 * d, bsv=Se * 2 * d, bso=S * d, scale=self.scale)
                ops.gemm(OC, st.shadow(ca.out_proj.weight), bias=ca.out_proj.bias, out=Y)
                ops.add_layernorm_fwd(XM, Y, XC, LN2, None, layer.norm2.weight, layer.norm2.bias, m2, r2, M, d,
                                      layer.norm2.eps)
                XM = XC
            else:
                ops.add_layernorm_fwd(XA, Y, XM, LN2, None, layer.norm2.weight, layer.norm2.bias, m2, r2, M, d,
                                      layer.norm2.eps)
            ops.gemm(LN2, st.shadow(layer.linear1.weight), bias=layer.linear1.bias, epilogue=ops.EPI_BF16_ACT, out=PRE,
                     out2=HACT, act=self.act)
            ops.gemm(HACT, st.shadow(layer.linear2.weight), bias=layer.linear2.bias, out=Y)
            XM_prev = XM
        self.saved = training
        self._X0 = X0 if training else None
        self._save.X0 = self._X0
        return XM_prev, Y

    def top_bias_grad(self) -> torch.Tensor:
        """Gradient slot of the last layer's linear2.bias: the producer of the incoming Gb sums its columns into it."""
        return self.store.grad(self.layers[-1].linear2.bias)

    def backward(self, G: torch.Tensor, Gb: torch.Tensor, B: int, S: int, on_layer_done=None,
                 top_bias_done: bool = False, save: Optional["Workspace"] = None) -> torch.Tensor:
        """G (fp32) / Gb (bf16 copy): gradient w.r.t. the final residual stream [B*S, d].  Returns G w.r.t. X0
        (in place).  Parameter gradients are ACCUMULATED into the ParamStore's flat fp32 buffer.
        The bias gradients of linear2 / out_proj are column sums of Gb; they are produced by the LayerNorm-backward
        kernel that writes Gb (`gsum`), not by a separate pass (top_bias_done: the caller's kernel did the top one)."""
        if save is None:
            if not self.saved:
                raise MMBError("backward called without a saved training forward")
            save = self.ws
        X0, kmask = save.X0, getattr(save, "kmask", None)
        mask3, enc, Se = getattr(save, "mask3", None), getattr(save, "enc", None), getattr(save, "S_enc", 0)
        save.dENC = None
        if X0 is None:
            raise MMBError("backward called without a saved training forward")
        st, d, ff, H = self.store, self.d, self.ff, self.H
        M = B * S
        bf = torch.bfloat16
        T1 = self.ws.get(f"{self.prefix}.T1", (M, d), bf)
        T3 = self.ws.get(f"{self.prefix}.T3", (M, 3 * d), bf)
        sp = ops.wgrad_splits
        for l in range(self.L - 1, -1, -1):
            layer = self.layers[l]
            at = layer.self_attn
            f32 = torch.float32
            g = lambda n, shape, dt: save.get(f"{self.prefix}.{n}.{l}", shape, dt)  # noqa: E731
            LN1, QKV, O = g("LN1", (M, d), bf), g("QKV", (M, 3 * d), bf), g("O", (M, d), bf)
            LSE = g("LSE", (B * H * S,), f32)
            XM, LN2 = g("XM", (M, d), f32), g("LN2", (M, d), bf)
            ca = getattr(layer, "cross_attn", None) if enc is not None else None
            XMID = g("XC", (M, d), f32) if ca is not None else XM     # the stream the MLP branch was added to
            PRE, HACT = g("PRE", (M, ff), bf), g("HACT", (M, ff), bf)
            m1, r1, m2, r2 = g("m1", (M,), f32), g("r1", (M,), f32), g("m2", (M,), f32), g("r2", (M,), f32)
            XA = X0 if l == 0 else g("XA", (M, d), f32)
            # ---- MLP branch:  y = W2 act(W1 LN2(x) + b1) + b2 ----
            ops.gemm(Gb, HACT, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(layer.linear2.weight),
                     splits=sp(d, ff, M), accumulate=True)
            if l == self.L - 1 and not top_bias_done:
                ops.colsum_bf16(Gb, st.grad(layer.linear2.bias), M, d, d)
            dPRE = HACT  # overwrite: act output is dead once its wgrad has been issued (same stream)
            fuse = os.environ.get("MMB_FUSE_COLSUM_GEMM", "1") == "1"
            ops.gemm(Gb, st.shadow(layer.linear2.weight), b_mn=True, epilogue=ops.EPI_BF16_DACT, aux=PRE, out=dPRE,
                     act=self.act, colsum=st.grad(layer.linear1.bias) if fuse else None)   # db1 = colsum(dPRE), fused into the epilogue
            ops.gemm(dPRE, LN2, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(layer.linear1.weight),
                     splits=sp(ff, d, M), accumulate=True)
            if not fuse:
                ops.colsum_bf16(dPRE, st.grad(layer.linear1.bias), M, ff, ff)
            ops.gemm(dPRE, st.shadow(layer.linear1.weight), b_mn=True, out=T1)
            ops.layernorm_bwd(XMID, T1, None, m2, r2, layer.norm2.weight, G, G, Gb, st.grad(layer.norm2.weight),
                              st.grad(layer.norm2.bias), M, d,
                              gsum=st.grad(ca.out_proj.bias if ca is not None else at.out_proj.bias))
            if ca is not None:
                # ---- cross-attention branch:  x2 = x1 + Wo_c Attn(Wq LN_c(x1), Wkv enc) ----
                lnc = layer.norm_cross
                dkv = enc.shape[1]
                LNC, QC, OC = g("LNC", (M, d), bf), g("QC", (M, d), bf), g("OC", (M, d), bf)
                KVC = g("KVC", (B * Se, 2 * d), bf)
                mc, rc = g("mc", (M,), f32), g("rc", (M,), f32)
                ops.gemm(Gb, OC, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(ca.out_proj.weight),
                         splits=sp(d, d, M), accumulate=True)
                ops.gemm(Gb, st.shadow(ca.out_proj.weight), b_mn=True, out=T1)      # d OC
                dQC = self.ws.get(f"{self.prefix}.dQC", (M, d), bf)
                dKVC = self.ws.get(f"{self.prefix}.dKVC", (B * Se, 2 * d), bf)
                ops.attention_bwd_generic(QC, KVC[:, :d], KVC[:, d:], T1, dKVC[:, :d], dKVC[:, d:], dq=dQC, B=B, Sq=S,
                                          Skv=Se, H=H, head_dim=64, bsq=S * d, bsk=Se * 2 * d, bsv=Se * 2 * d, bso=S * d,
                                          scale=self.scale)
                ops.gemm(dQC, LNC, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(ca.q_w), splits=sp(d, d, M),
                         accumulate=True)
                ops.colsum_bf16(dQC, st.grad(ca.q_b), M, d, d)
                ops.gemm(dKVC, enc, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(ca.kv_w),
                         splits=sp(2 * d, dkv, B * Se), accumulate=True)
                ops.colsum_bf16(dKVC, st.grad(ca.kv_b), B * Se, 2 * d, 2 * d)
                if save.dENC is None:   # gradient w.r.t. the cross-attention source, summed over the layers
                    save.dENC = torch.empty((B * Se, dkv), device=G.device, dtype=f32)
                    ops.gemm(dKVC, st.shadow(ca.kv_w), b_mn=True, epilogue=ops.EPI_F32, out=save.dENC)
                else:
                    ops.gemm(dKVC, st.shadow(ca.kv_w), b_mn=True, epilogue=ops.EPI_F32, out=save.dENC, accumulate=True)
                ops.gemm(dQC, st.shadow(ca.q_w), b_mn=True, out=T1)                  # d LN_c(x1)
                ops.layernorm_bwd(XM, T1, None, mc, rc, lnc.weight, G, G, Gb, st.grad(lnc.weight), st.grad(lnc.bias), M, d,
                                  gsum=st.grad(at.out_proj.bias))
            # ---- attention branch ----
            ops.gemm(Gb, O, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(at.out_proj.weight),
                     splits=sp(d, d, M), accumulate=True)
            ops.gemm(Gb, st.shadow(at.out_proj.weight), b_mn=True, out=T1)  # dO
            if mask3 is not None:
                ops.attention_bwd_generic(QKV[:, :d], QKV[:, d:2 * d], QKV[:, 2 * d:], T1, T3[:, d:2 * d], T3[:, 2 * d:],
                                          dq=T3[:, :d], B=B, Sq=S, Skv=S, H=H, head_dim=64, bsq=S * 3 * d, bsk=S * 3 * d,
                                          bsv=S * 3 * d, bso=S * d, scale=self.scale, mask=mask3, mask_bs=S * S, mask_qs=S,
                                          causal=self.causal)
            elif kmask is not None:
                ops.attention_bwd_kmask(QKV, O, T1, LSE, T3, kmask, B, S, H, self.causal, self.scale)
            else:
                ops.attention_bwd(QKV, O, T1, LSE, T3, B, S, H, self.causal, self.scale)
            ops.gemm(T3, LN1, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(at.in_proj_weight),
                     splits=sp(3 * d, d, M), accumulate=True)
            ops.colsum_bf16(T3, st.grad(at.in_proj_bias), M, 3 * d, 3 * d)
            ops.gemm(T3, st.shadow(at.in_proj_weight), b_mn=True, out=T1)
            ops.layernorm_bwd(XA, T1, None, m1, r1, layer.norm1.weight, G, G, Gb, st.grad(layer.norm1.weight),
                              st.grad(layer.norm1.bias), M, d,
                              gsum=st.grad(self.layers[l - 1].linear2.bias) if l > 0 else None)
            if on_layer_done is not None:
                on_layer_done(l)  # all parameter gradients of layer l are final (data-parallel all-reduce hook)
        return G


class ViTTower:
    """CLIPViTEncoder runtime (models/clip/image_encoder.py:82-113)."""

    def __init__(self, mod: nn.Module):
        self.mod = mod
        self.store = ParamStore(list(mod.parameters()))
        self.ws = Workspace(self.store.device)
        d = mod.conv.weight.shape[0]
        layer0 = mod.encoder.layers[0]
        self.d, self.ps = d, mod.conv.weight.shape[2]
        self.E = mod.projection.shape[1]
        self.stack = TransformerStack(mod.encoder.layers, self.store, self.ws, d=d, heads=layer0.self_attn.num_heads,
                                      ff=layer0.linear1.weight.shape[0], causal=False, act=ops.ACT_QUICK_GELU,
                                      prefix="img")
        self.gen = 0

    def forward(self, image: torch.Tensor, training: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """out (optional): fp32 [B, E] destination of the embeddings (e.g. a row block of a larger buffer)."""
        mod, st, ws, d = self.mod, self.store, self.ws, self.d
        if image.dtype != torch.float32:
            image = image.float()
        image = image.contiguous()
        B, _, Himg, Wimg = image.shape
        ps = self.ps
        P = (Himg // ps) * (Wimg // ps)
        S = P + 1
        K = 3 * ps * ps
        bf, f32 = torch.bfloat16, torch.float32
        st.refresh()
        Kp = -(-K // 8) * 8   # row pitch: bf16 rows must be 16 B multiples for TMA (K = 588 -> 592 for 14x14 patches)
        PATCH = ws.get("img.PATCH", (B * P, Kp), bf)[:, :K]
        PO = ws.get("img.PO", (B * P, d), bf)
        X0 = ws.get("img.X0", (B * S, d), f32)
        m0 = ws.get("img.m0", (B * S,), f32); r0 = ws.get("img.r0", (B * S,), f32)
        ops.im2col(image, ps, PATCH)
        wconv = st.shadow2d(mod.conv.weight)
        if Kp != K:  # re-pitch the (tiny) conv weight shadow the same way
            wpad = ws.get("img.WCONV", (d, Kp), bf)[:, :K]
            wpad.copy_(wconv)
            wconv = wpad
        ops.gemm(PATCH, wconv, out=PO)
        ops.vit_embed_ln_fwd(PO, mod.cls_token_embedding, mod.positional_embedding, mod.ln_pre.weight, mod.ln_pre.bias,
                             X0, m0, r0, B, S, d, mod.ln_pre.eps)
        XM, Y = self.stack.forward(X0, B, S, training)
        XSEL = ws.get("img.XSEL", (B, d), f32)
        LNP = ws.get("img.LNP", (B, d), bf)
        mP = ws.get("img.mP", (B,), f32); rP = ws.get("img.rP", (B,), f32)
        ops.add_layernorm_fwd(XM, Y, XSEL, LNP, None, mod.ln_post.weight, mod.ln_post.bias, mP, rP, B, d, mod.ln_post.eps,
                              row_idx=None, rows_per_group=S)
        EMB = out if out is not None else torch.empty((B, self.E), device=image.device, dtype=f32)
        ops.gemm(LNP, st.shadow(mod.projection), b_mn=True, epilogue=ops.EPI_F32, out=EMB)
        self.B, self.S, self.P = B, S, P
        self.gen += 1
        return EMB

    def backward(self, dEMB: torch.Tensor) -> None:
        mod, st, ws, d = self.mod, self.store, self.ws, self.d
        B, S, P = self.B, self.S, self.P
        bf, f32 = torch.bfloat16, torch.float32
        M = B * S
        dEb = ops.cast_bf16(dEMB.contiguous())
        LNP, XSEL = ws.get("img.LNP", (B, d), bf), ws.get("img.XSEL", (B, d), f32)
        ops.gemm(LNP, dEb, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(mod.projection), accumulate=True)
        dLNP = ws.get("img.dLNP", (B, d), f32)
        ops.gemm(dEb, st.shadow(mod.projection), epilogue=ops.EPI_F32, out=dLNP)
        G = ws.get("img.G", (M, d), f32)
        Gb = ws.get("img.Gb", (M, d), bf)
        ops.zero_(G); ops.zero_(Gb)
        ops.layernorm_bwd(XSEL, None, dLNP, ws.get("img.mP", (B,), f32), ws.get("img.rP", (B,), f32), mod.ln_post.weight, None, G, Gb,
                          st.grad(mod.ln_post.weight), st.grad(mod.ln_post.bias), B, d, row_idx=None, rows_per_group=S,
                          gsum=self.stack.top_bias_grad())
        self.stack.backward(G, Gb, B, S, on_layer_done=getattr(self, "layer_done_cb", None), top_bias_done=True)
        PO = ws.get("img.PO", (B * P, d), bf)
        DP = ws.get("img.DP", (B * P, d), bf)
        ops.vit_embed_ln_bwd(PO, mod.cls_token_embedding, mod.positional_embedding, G, ws.get("img.m0", (M,), f32), ws.get("img.r0", (M,), f32),
                             mod.ln_pre.weight, G, DP, st.grad(mod.ln_pre.weight), st.grad(mod.ln_pre.bias), B, S, d)
        ops.batch_sum(G, st.grad(mod.positional_embedding), B, S * d, S * d)
        ops.batch_sum(G, st.grad(mod.cls_token_embedding), B, S * d, d)
        K = 3 * self.ps * self.ps
        PATCH = ws.get("img.PATCH", (B * P, -(-K // 8) * 8), bf)[:, :K]
        ops.gemm(DP, PATCH, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad2d(mod.conv.weight),
                 splits=ops.wgrad_splits(d, PATCH.shape[1], B * P), accumulate=True)


class TextTower:
    """CLIPTextEncoder runtime (models/clip/text_encoder.py:113-134)."""

    def __init__(self, mod: nn.Module):
        self.mod = mod
        self.store = ParamStore(list(mod.parameters()))
        self.ws = Workspace(self.store.device)
        layer0 = mod.encoder.layers[0]
        self.d = mod.width
        self.E = mod.projection.weight.shape[0]
        self.stack = TransformerStack(mod.encoder.layers, self.store, self.ws, d=self.d,
                                      heads=layer0.self_attn.num_heads, ff=layer0.linear1.weight.shape[0], causal=True,
                                      act=ops.ACT_QUICK_GELU, prefix="txt")
        self.gen = 0

    def forward(self, text: torch.Tensor, training: bool, return_hidden_state: bool = False,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        mod, st, ws, d = self.mod, self.store, self.ws, self.d
        if text.dtype != torch.int64:
            text = text.long()
        text = text.contiguous()
        B, S = text.shape
        bf, f32 = torch.bfloat16, torch.float32
        st.refresh()
        X0 = ws.get("txt.X0", (B * S, d), f32)
        V = mod.token_embedding.weight.shape[0]
        ops.text_embed_fwd(text, mod.token_embedding.weight, mod.positional_embedding, X0, B, S, d, V)
        XM, Y = self.stack.forward(X0, B, S, training and not return_hidden_state)
        if return_hidden_state:
            HS = torch.empty((B, S, d), device=text.device, dtype=f32)
            ops.add_layernorm_fwd(XM, Y, None, None, HS, mod.ln_final.weight, mod.ln_final.bias, None, None, B * S, d,
                                  mod.ln_final.eps)
            return HS
        IDX = ws.get("txt.IDX", (B,), torch.int32)
        ops.argmax_tokens(text, IDX, B, S)
        XSEL = ws.get("txt.XSEL", (B, d), f32)
        LNF = ws.get("txt.LNF", (B, d), bf)
        mF = ws.get("txt.mF", (B,), f32); rF = ws.get("txt.rF", (B,), f32)
        ops.add_layernorm_fwd(XM, Y, XSEL, LNF, None, mod.ln_final.weight, mod.ln_final.bias, mF, rF, B, d,
                              mod.ln_final.eps, row_idx=IDX, rows_per_group=S)
        EMB = out if out is not None else torch.empty((B, self.E), device=text.device, dtype=f32)
        ops.gemm(LNF, st.shadow(mod.projection.weight), epilogue=ops.EPI_F32, out=EMB)
        self.B, self.S = B, S
        self.tokens = text if training else None
        self.gen += 1
        return EMB

    def backward(self, dEMB: torch.Tensor) -> None:
        mod, st, ws, d = self.mod, self.store, self.ws, self.d
        B, S = self.B, self.S
        bf, f32 = torch.bfloat16, torch.float32
        M = B * S
        dEb = ops.cast_bf16(dEMB.contiguous())
        LNF, XSEL = ws.get("txt.LNF", (B, d), bf), ws.get("txt.XSEL", (B, d), f32)
        ops.gemm(dEb, LNF, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=st.grad(mod.projection.weight), accumulate=True)
        dLNF = ws.get("txt.dLNF", (B, d), f32)
        ops.gemm(dEb, st.shadow(mod.projection.weight), b_mn=True, epilogue=ops.EPI_F32, out=dLNF)
        G = ws.get("txt.G", (M, d), f32)
        Gb = ws.get("txt.Gb", (M, d), bf)
        ops.zero_(G); ops.zero_(Gb)
        IDX = ws.get("txt.IDX", (B,), torch.int32)
        ops.layernorm_bwd(XSEL, None, dLNF, ws.get("txt.mF", (B,), f32), ws.get("txt.rF", (B,), f32), mod.ln_final.weight, None, G, Gb,
                          st.grad(mod.ln_final.weight), st.grad(mod.ln_final.bias), B, d, row_idx=IDX, rows_per_group=S,
                          gsum=self.stack.top_bias_grad())
        self.stack.backward(G, Gb, B, S, top_bias_done=True)
        ops.batch_sum(G, st.grad(mod.positional_embedding), B, S * d, S * d)
        ops.text_embed_bwd(self.tokens, G, st.grad(mod.token_embedding.weight), B, S, d)
