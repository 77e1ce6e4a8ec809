"""Parameter-free activation marker modules (mirror of torchmultimodal/modules/layers/activation.py:12-25).

``SiLU`` is the reference's (mis)named QuickGELU, ``sigmoid(1.702 x) * x``.  In this package it only TAGS which fused
GEMM epilogue a transformer stack uses (``multimodal_b200.ops.ACT_QUICK_GELU``); the arithmetic lives in
``csrc/gemm.cu``.  Called on its own on a CUDA tensor it runs one elementwise kernel (forward values).
"""
from torch import nn, Tensor

from ..._lib import MMBError


class SiLU(nn.Module):
    r"""QuickGELU marker: :math:`x \cdot \sigma(1.702 x)` (computed inside the FC1 GEMM epilogue)."""

    def forward(self, x: Tensor) -> Tensor:
        """Standalone: one elementwise kernel (values only; inside the encoders it is the FC1 GEMM epilogue)."""
        import torch

        from ... import ops

        if torch.is_grad_enabled() and x.requires_grad:
            raise MMBError("standalone SiLU (QuickGELU) has no autograd path; it is fused inside the encoder runtimes")
        if not x.is_cuda:
            raise MMBError("SiLU: expected a CUDA tensor (multimodal_b200 has no CPU path)")
        return ops.act_fwd(x.contiguous().float(), ops.ACT_QUICK_GELU).to(x.dtype)
