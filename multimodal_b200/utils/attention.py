"""Mirror of torchmultimodal/utils/attention.py:13-65."""
from typing import Optional

import torch
from torch import Tensor


def get_extended_attention_mask(attention_mask: Tensor) -> Tensor:
    if attention_mask.dim() == 4:
        extended_attention_mask = attention_mask
    elif attention_mask.dim() == 3:
        extended_attention_mask = attention_mask[:, None, :, :]
    elif attention_mask.dim() == 2:
        extended_attention_mask = attention_mask[:, None, None, :]
    else:
        raise ValueError("Wrong shape for attention_mask (shape {})".format(attention_mask.shape))
    return extended_attention_mask.to(dtype=attention_mask.dtype)


def get_causal_attention_mask(tgt_seq_len: int, src_seq_len: Optional[int] = None) -> Tensor:
    """Lower-triangular ones of shape (tgt_seq_len, src_seq_len) (utils/attention.py:56-65)."""
    if src_seq_len is None:
        src_seq_len = tgt_seq_len
    return torch.tril(torch.ones(tgt_seq_len, src_seq_len))
