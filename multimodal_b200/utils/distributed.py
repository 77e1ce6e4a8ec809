"""Mirror of torchmultimodal/utils/distributed.py:16-90 (the library's only communication entry points).

``gather_tensor`` keeps the reference's three autograd modes.  On the contrastive-loss hot path it is NOT used: the
loss pulls peer embeddings inside its own kernel (see modules/losses/contrastive_loss_with_temperature.py).  It stays
available for callers that gather other tensors, implemented with torch.distributed (plumbing).
"""
from enum import Enum
from typing import List

import torch
from torch import Tensor
from torch.distributed import all_gather as all_gather_no_backprop
from torch.distributed.nn.functional import all_gather as all_gather_with_backprop


class BackpropType(Enum):
    """GLOBAL: gradients flow to every worker; LOCAL: only to the local slot; NONE: no gradient."""

    GLOBAL = 0
    LOCAL = 1
    NONE = 2


def get_rank() -> int:
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return 0


def gather_tensor(tensor: Tensor, backprop_type: BackpropType = BackpropType.GLOBAL) -> List[Tensor]:
    world_size = torch.distributed.get_world_size()
    if backprop_type == BackpropType.GLOBAL:
        return list(all_gather_with_backprop(tensor))
    tensor_all_gpus = [torch.zeros_like(tensor) for _ in range(world_size)]
    all_gather_no_backprop(tensor_all_gpus, tensor)
    if backprop_type == BackpropType.LOCAL:
        tensor_all_gpus[get_rank()] = tensor
    return tensor_all_gpus


def concat_gather_all_gpu(tensor: Tensor, backprop_type: BackpropType = BackpropType.GLOBAL, dim: int = 0) -> Tensor:
    if not torch.distributed.is_available() or not torch.distributed.is_initialized():
        return tensor
    return torch.cat(gather_tensor(tensor, backprop_type), dim=dim)
