"""torch.autograd glue: lets the explicit forward/backward runtimes (engine.py) sit behind ordinary nn.Modules so
that ``model(x); loss.backward(); torch.optim...step()`` works exactly as with the reference modules."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import MMBError


class TowerFunction(torch.autograd.Function):
    """Whole-encoder forward/backward.  inputs: (runtime, data, *parameters)."""

    @staticmethod
    def forward(ctx, runtime, data, *params):
        training = any(ctx.needs_input_grad[2:])
        emb = runtime.forward(data, training)
        ctx.runtime, ctx.gen, ctx.n = runtime, runtime.gen, len(params)
        ctx.need = ctx.needs_input_grad[2:]
        return emb

    @staticmethod
    def backward(ctx, demb):
        rt = ctx.runtime
        if rt.gen != ctx.gen:
            raise MMBError("the encoder ran another forward before this backward: its saved activations were "
                           "overwritten (one in-flight training forward per encoder instance)")
        st = rt.store
        flat = st.master is not None
        if not flat:
            st.zero_grads()
        rt.backward(demb.contiguous().float())
        if flat:  # p.grad are views of the flat buffer: gradients were accumulated in place
            # `optimizer.zero_grad(set_to_none=True)` (torch's default) or `p.grad = None` severs those views; re-attach
            # them, otherwise gradients would pile up invisibly in the flat buffer while the optimizer skips the parameter
            for p in st.params:
                want = st.grad(p)
                if p.grad is None or p.grad.data_ptr() != want.data_ptr():
                    p.grad = want
            return (None, None) + (None,) * ctx.n
        g = st.g.clone()
        grads = []
        for p, need in zip(st.params, ctx.need):
            o = st.off[id(p)]
            grads.append(g[o:o + p.numel()].view(p.shape) if need else None)
        return (None, None, *grads)


class L2NormalizeFunction(torch.autograd.Function):
    """F.normalize(x) along dim=1 for 2-D inputs (models/clip/model.py:72-73)."""

    @staticmethod
    def forward(ctx, x, eps):
        if x.dim() != 2:
            raise MMBError("L2 normalise kernel expects a 2-D [batch, embedding] tensor")
        xf = x.contiguous().float()
        B, E = xf.shape
        y = torch.empty_like(xf)
        inv = torch.empty(B, device=x.device, dtype=torch.float32)
        ops.l2norm_fwd(xf, y, None, inv, B, E, eps)
        ctx.save_for_backward(y, inv)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        B, E = y.shape
        dx = torch.empty_like(y)
        ops.l2norm_bwd(dy.contiguous().float(), y, inv, dx, None, B, E)
        return dx.to(dy.dtype), None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    return L2NormalizeFunction.apply(x, eps)


def autocast_out(t: torch.Tensor) -> torch.Tensor:
    """Dtype convention of the reference under ``torch.autocast`` (examples/flava/native/train.py:296-298): the towers'
    last op is a matmul / Linear, whose autocast output is the autocast dtype (bf16).  The kernels always produce fp32
    embeddings; inside an autocast region they are cast at the module boundary (differentiably), outside they stay fp32
    as the reference's fp32 modules return."""
    if t.is_cuda and torch.is_autocast_enabled():
        return t.to(torch.get_autocast_gpu_dtype())
    return t
