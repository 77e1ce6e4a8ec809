"""`Pooler` — mirror of torchmultimodal/modules/losses/flava.py:84-97 (dense + tanh on the first token), the only
piece of that file on the encoder path.  Parameter container; executed inside the encoder runtimes."""
from typing import Any

from torch import nn, Tensor

from ..._lib import MMBError


class Pooler(nn.Module):
    def __init__(self, hidden_size: int = 768, **kwargs: Any):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states: Tensor) -> Tensor:
        raise MMBError("Pooler is fused into the encoder runtime; not a standalone op here")
