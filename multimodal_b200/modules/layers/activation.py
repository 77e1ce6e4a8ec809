"""Parameter-free activation marker modules (mirror of torchmultimodal/modules/layers/activation.py:12-25).

``SiLU`` is the reference's (mis)named QuickGELU, ``sigmoid(1.702 x) * x``.  In this package it only TAGS which fused
GEMM epilogue a transformer stack uses (``multimodal_b200.ops.ACT_QUICK_GELU``); the arithmetic lives in
``csrc/gemm.cu``.  Calling it on a CUDA tensor is not part of the hot path and is therefore not provided.
"""
from torch import nn, Tensor

from ..._lib import MMBError


class SiLU(nn.Module):
    r"""QuickGELU marker: :math:`x \cdot \sigma(1.702 x)` (computed inside the FC1 GEMM epilogue)."""

    def forward(self, x: Tensor) -> Tensor:
        raise MMBError("SiLU (QuickGELU) is fused into the MLP GEMM epilogue; it is not a standalone op here")
