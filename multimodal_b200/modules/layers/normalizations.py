"""Mirror of torchmultimodal/modules/layers/normalizations.py:13-25 (Fp32LayerNorm).

A parameter container with the reference's state-dict keys (``weight``, ``bias``).  Inside the encoders the
normalisation is executed by ``mmb_add_layernorm_fwd`` (fp32 statistics — which IS the Fp32LayerNorm contract);
called standalone on a CUDA tensor it runs the same kernel.
"""
import torch
from torch import nn, Tensor

from ... import ops
from ..._lib import MMBError


class Fp32LayerNorm(nn.LayerNorm):
    def forward(self, x: Tensor) -> Tensor:
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            raise MMBError("standalone Fp32LayerNorm has no autograd path; it is fused inside the encoder runtimes")
        d = x.shape[-1]
        xf = x.float().contiguous().view(-1, d)
        out = torch.empty_like(xf)
        ops.add_layernorm_fwd(xf, None, None, None, out, self.weight, self.bias, None, None, xf.shape[0], d, self.eps)
        return out.view(x.shape).type_as(x)
