"""Mirror of torchmultimodal/utils/attention.py:13-53."""
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
