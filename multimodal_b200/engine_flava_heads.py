"""Backward schedules of the FLAVA pre-training heads as torch.autograd Functions over the library's
kernels, so that `FLAVAPretrainingLoss` / `FLAVAForPreTraining` train with ``loss.backward()`` like the reference
(torchmultimodal/modules/losses/flava.py:84-484 under autograd).

* ``MaskedPredictionFunction`` — the whole MLM / MIM / MMM head in one node: boolean-mask row select -> dense GEMM +
  bias + erf-GELU -> fp32 LayerNorm -> vocabulary GEMM + bias -> mean cross-entropy(ignore_index).  Backward:
  d logits (bf16) from the saved logits (`mmb_ce_labels_bwd`, scaled by the incoming d loss on the device), decoder
  weight / bias gradients (wgrad GEMM, column sum), LayerNorm backward, GELU' (`mmb_act_bwd`), dense weight / bias
  gradients, and the row gradients scatter-added into a dense [B, S, d] gradient (`mmb_scatter_rows_idx_add`).
* ``SmallLinearFunction`` — Linear with fewer than 8 outputs (TwoWayHead, losses/flava.py:100-110): tensor-core forward
  (outputs padded to 8), exact fp32 SIMT backward (`mmb_matmul_f32`).
* ``CrossEntropyFunction`` — `nn.CrossEntropyLoss(ignore_index)` on materialised fp32 logits (ITM, :113-140).
The Pooler and the global contrastive loss reuse ``engine_flava_train.FirstTokenLinearFunction``,
``autograd.L2NormalizeFunction`` and ``engine_loss.ContrastiveFunction`` (all have backward schedules already).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import MMBError


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _rows(x: torch.Tensor) -> torch.Tensor:
    """[B, S, d] fp32 view whose rows are d contiguous floats and whose batch pitch is a multiple of the row pitch."""
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(-1) != 1 or (x.dim() == 3 and x.stride(0) % x.stride(1) != 0):
        x = x.contiguous()
    return x


def _bf16_weight(w: torch.Tensor, rows: Optional[int] = None) -> torch.Tensor:
    """bf16 copy of a [N, K] weight, zero-padded to `rows` rows (TMA row counts / pitches in multiples of 8)."""
    wb = ops.cast_bf16(w.detach().contiguous())
    if rows is not None and rows != w.shape[0]:
        wp = torch.zeros((rows, w.shape[1]), device=w.device, dtype=torch.bfloat16)
        wp[:w.shape[0]].copy_(wb)
        wb = wp
    return wb


def _f32_bias(b: Optional[torch.Tensor], n: int) -> Optional[torch.Tensor]:
    if b is None:
        return None
    bf = b.detach().float().contiguous()
    if n != bf.numel():
        bp = torch.zeros(n, device=b.device, dtype=torch.float32)
        bp[:bf.numel()].copy_(bf)
        bf = bp
    return bf


class MaskedPredictionFunction(torch.autograd.Function):
    """(logits [n, V] fp32 (not differentiable), loss) = head(hidden[keep]) against `labels` [n] (all kept rows)."""

    @staticmethod
    def forward(ctx, hidden, keep, labels, ignore_index, eps, dense_w, dense_b, ln_w, ln_b, dec_w, dec_b):
        ctx.set_materialize_grads(False)
        x = _rows(hidden)
        B, S, d = x.shape
        dev = x.device
        bf, f32 = torch.bfloat16, torch.float32
        bs = keep.nonzero(as_tuple=False)                      # [n, 2] (b, s): bookkeeping on the labels
        n = int(bs.shape[0])
        V = dec_w.shape[0]
        Vp = _pad8(V)
        ctx.meta = (tuple(hidden.shape), n, V, Vp, d, int(ignore_index), float(eps), hidden.dtype)
        if n == 0:   # CE mean over zero rows (NaN, as torch); nothing to back-propagate
            ctx.save_for_backward()
            return torch.empty((0, V), device=dev, dtype=f32), torch.full((), float("nan"), device=dev)
        ld = x.stride(1)
        idx = (bs[:, 0] * (x.stride(0) // ld) + bs[:, 1]).contiguous()
        rows = torch.empty((n, d), device=dev, dtype=bf)
        ops.gather_rows_idx_cast(x, idx, rows, d)
        wd = _bf16_weight(dense_w)
        pre, act = ops.gemm(rows, wd, bias=_f32_bias(dense_b, d), epilogue=ops.EPI_BF16_ACT, act=ops.ACT_GELU_ERF)
        xact = torch.empty((n, d), device=dev, dtype=f32)       # fp32 LayerNorm input, kept for the backward
        ln = torch.empty((n, d), device=dev, dtype=bf)
        mean, rstd = torch.empty(n, device=dev, dtype=f32), torch.empty(n, device=dev, dtype=f32)
        lw = ln_w.detach().float().contiguous()
        ops.add_layernorm_fwd(None, act, xact, ln, None, lw, ln_b.detach().float().contiguous(), mean, rstd, n, d, eps)
        wdec = _bf16_weight(dec_w, Vp)
        logits = torch.empty((n, Vp), device=dev, dtype=f32)
        ops.gemm(ln, wdec, epilogue=ops.EPI_F32, bias=_f32_bias(dec_b, Vp), out=logits)
        lab = labels.contiguous().long()
        accum = torch.zeros(2, device=dev, dtype=f32)
        ops.ce_labels(logits[:, :V], lab, 1, ignore_index, n, V, None, accum)
        ctx.save_for_backward(rows, pre, xact, mean, rstd, ln, logits, lab, accum, wd, wdec, lw, bs)
        out_logits = logits[:, :V]
        ctx.mark_non_differentiable(out_logits)
        return out_logits, accum[0] / accum[1]

    @staticmethod
    def backward(ctx, _dlogits, dloss):
        shape, n, V, Vp, d, ignore_index, eps, in_dtype = ctx.meta
        need = ctx.needs_input_grad
        none = (None,) * 11
        if n == 0 or dloss is None:
            return none
        rows, pre, xact, mean, rstd, ln, logits, lab, accum, wd, wdec, lw, bs = ctx.saved_tensors
        dev = rows.device
        bf, f32 = torch.bfloat16, torch.float32
        gs = dloss.detach().float().reshape(1).contiguous()
        dlog = torch.zeros((n, Vp), device=dev, dtype=bf)       # pad columns stay zero
        ops.ce_labels_bwd(logits[:, :V], lab, 1, ignore_index, n, V, accum, 1.0, dlog[:, :V], gscale=gs)
        d_dec_w = d_dec_b = d_dense_w = d_dense_b = d_ln_w = d_ln_b = d_hidden = None
        if need[9]:
            gw = torch.empty((Vp, d), device=dev, dtype=f32)
            ops.gemm(dlog, ln, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=gw, splits=ops.wgrad_splits(Vp, d, n))
            d_dec_w = gw[:V]
        if need[10]:
            gb = torch.zeros(Vp, device=dev, dtype=f32)
            ops.colsum_bf16(dlog, gb, n, Vp, Vp)
            d_dec_b = gb[:V]
        dln = torch.empty((n, d), device=dev, dtype=bf)
        ops.gemm(dlog, wdec, b_mn=True, out=dln)
        dact = torch.empty((n, d), device=dev, dtype=bf)
        d_ln_w = torch.zeros(d, device=dev, dtype=f32)
        d_ln_b = torch.zeros(d, device=dev, dtype=f32)
        ops.layernorm_bwd(xact, dln, None, mean, rstd, lw, None, None, dact, d_ln_w, d_ln_b, n, d)
        dpre = torch.empty((n, d), device=dev, dtype=bf)
        ops.act_bwd(dact, pre, dpre, ops.ACT_GELU_ERF)
        if need[5]:
            d_dense_w = torch.empty((d, d), device=dev, dtype=f32)
            ops.gemm(dpre, rows, a_mn=True, b_mn=True, epilogue=ops.EPI_F32, out=d_dense_w,
                     splits=ops.wgrad_splits(d, d, n))
        if need[6]:
            d_dense_b = torch.zeros(d, device=dev, dtype=f32)
            ops.colsum_bf16(dpre, d_dense_b, n, d, d)
        if need[0]:
            drows = torch.empty((n, d), device=dev, dtype=f32)
            ops.gemm(dpre, wd, b_mn=True, epilogue=ops.EPI_F32, out=drows)
            d_hidden = torch.zeros(shape, device=dev, dtype=f32)
            idx = (bs[:, 0] * shape[1] + bs[:, 1]).contiguous()
            ops.scatter_rows_idx_add(drows, idx, d_hidden.view(-1, d), d)
            d_hidden = d_hidden.to(in_dtype)
        return (d_hidden, None, None, None, None, d_dense_w, d_dense_b, d_ln_w if need[7] else None,
                d_ln_b if need[8] else None, d_dec_w, d_dec_b)


class SmallLinearFunction(torch.autograd.Function):
    """y = x @ W^T + b for a 2-D fp32 x and fewer than 8 outputs."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        N, K = weight.shape
        Np = _pad8(N)
        xf = x.detach().float().contiguous()
        out = torch.empty((xf.shape[0], Np), device=x.device, dtype=torch.float32)
        ops.gemm(ops.cast_bf16(xf), _bf16_weight(weight, Np), epilogue=ops.EPI_F32, bias=_f32_bias(bias, Np), out=out)
        ctx.save_for_backward(xf, weight.detach().float().contiguous())
        ctx.has_bias = bias is not None
        return out[:, :N]

    @staticmethod
    def backward(ctx, dy):
        xf, w = ctx.saved_tensors
        dyf = dy.contiguous().float()
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.matmul_f32(dyf, w)                                   # [M, N] @ [N, K]
        if ctx.needs_input_grad[1]:
            dW = ops.matmul_f32(dyf, xf, ta=True)                         # [N, M] @ [M, K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            ones = torch.ones((1, dyf.shape[0]), device=dy.device, dtype=torch.float32)
            db = ops.matmul_f32(ones, dyf).view(-1)                       # column sums of dy
        return dx, dW, db


class CrossEntropyFunction(torch.autograd.Function):
    """nn.CrossEntropyLoss(ignore_index)(logits [M, V] fp32, labels [M]) — mean over the kept rows."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        lg = logits.detach().float()
        if lg.stride(-1) != 1:
            lg = lg.contiguous()
        M, V = lg.shape
        lab = labels.contiguous().long()
        accum = torch.zeros(2, device=lg.device, dtype=torch.float32)
        ops.ce_labels(lg, lab, 1, ignore_index, M, V, None, accum)
        ctx.save_for_backward(lg, lab, accum)
        ctx.ignore_index = int(ignore_index)
        return accum[0] / accum[1]

    @staticmethod
    def backward(ctx, dloss):
        lg, lab, accum = ctx.saved_tensors
        M, V = lg.shape
        Vp = _pad8(V)
        dl = torch.zeros((M, Vp), device=lg.device, dtype=torch.bfloat16)
        ops.ce_labels_bwd(lg, lab, 1, ctx.ignore_index, M, V, accum, 1.0, dl[:, :V],
                          gscale=dloss.detach().float().reshape(1).contiguous())
        out = torch.empty((M, Vp), device=lg.device, dtype=torch.float32)
        ops.cast_f32(dl, out)
        return out[:, :V], None, None


def masked_prediction(hidden, keep, labels, head, ignore_index):
    """-> (logits [n, V], mean cross-entropy) of `head` (a MaskedPredictionHead) on hidden[keep]."""
    if not isinstance(head.layer_norm, torch.nn.LayerNorm):
        raise MMBError("MaskedPredictionHead.layer_norm must be a LayerNorm")
    return MaskedPredictionFunction.apply(hidden, keep, labels, ignore_index, head.layer_norm.eps, head.dense.weight,
                                          head.dense.bias, head.layer_norm.weight, head.layer_norm.bias,
                                          head.decoder.weight, head.bias)


def small_linear(x, linear):
    return SmallLinearFunction.apply(x, linear.weight, linear.bias)


def cross_entropy(logits, labels, ignore_index):
    return CrossEntropyFunction.apply(logits, labels, ignore_index)
