"""CLIP container + builders — drop-in for torchmultimodal/models/clip/model.py:19-115 (ViT variants).

``clip_vit_b16`` / ``clip_vit_b32`` / ``clip_vit_l14`` build the same architectures with the same state-dict schema,
so reference checkpoints load with ``load_state_dict`` unchanged.  The ResNet variants (model.py:117-215) are outside
the hot path named by BASELINE.json and are not provided.
"""
from typing import NamedTuple

import torch
from torch import nn

from ...autograd import l2_normalize
from .image_encoder import CLIPViTEncoder
from .text_encoder import CLIPTextEncoder


class CLIPOutput(NamedTuple):
    embeddings_a: torch.Tensor
    embeddings_b: torch.Tensor


class CLIP(nn.Module):
    """Two encoders + L2 normalisation (model.py:36-74).  Any nn.Module pair producing [B, E] CUDA tensors works."""

    def __init__(self, encoder_a: nn.Module, encoder_b: nn.Module):
        super().__init__()
        self.encoder_a = encoder_a
        self.encoder_b = encoder_b

    def forward(self, features_a: torch.Tensor, features_b: torch.Tensor) -> CLIPOutput:
        embeddings_a = self.encoder_a(features_a)
        embeddings_b = self.encoder_b(features_b)
        embeddings_a = l2_normalize(embeddings_a)
        embeddings_b = l2_normalize(embeddings_b)
        return CLIPOutput(embeddings_a=embeddings_a, embeddings_b=embeddings_b)


def _no_pretrained(pretrained: bool) -> None:
    if pretrained:
        raise RuntimeError("pretrained=True needs network access (reference: utils/common.py:99-108); load a "
                           "reference-format checkpoint with load_state_dict instead — the key schema is identical")


def clip_vit_b16(pretrained: bool = False) -> CLIP:
    _no_pretrained(pretrained)
    vision_encoder = CLIPViTEncoder(image_size=224, patch_size=16, layers=12, heads=12, width=768, embedding_dim=512)
    text_encoder = CLIPTextEncoder(embedding_dim=512)
    return CLIP(vision_encoder, text_encoder)


def clip_vit_b32(pretrained: bool = False) -> CLIP:
    _no_pretrained(pretrained)
    vision_encoder = CLIPViTEncoder(image_size=224, patch_size=32, layers=12, heads=12, width=768, embedding_dim=512)
    text_encoder = CLIPTextEncoder(embedding_dim=512)
    return CLIP(vision_encoder, text_encoder)


def clip_vit_l14(pretrained: bool = False) -> CLIP:
    _no_pretrained(pretrained)
    vision_encoder = CLIPViTEncoder(image_size=224, patch_size=14, layers=24, heads=16, width=1024, embedding_dim=768)
    text_encoder = CLIPTextEncoder(embedding_dim=768, width=768, dim_feedforward=3072, heads=12)
    return CLIP(vision_encoder, text_encoder)
