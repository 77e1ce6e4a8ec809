"""Mirror of torchmultimodal/modules/layers/transformer.py: `TransformerOutput` (:22-28, the NamedTuple every FLAVA /
CoCa encoder returns) and the parameter containers `TransformerEncoderLayer` / `TransformerEncoder` (:31-259) and
`TransformerDecoderLayer` / `TransformerDecoder` (:262-657) — same constructors, state-dict keys and creation order.
The layers execute inside `engine_coca.LayerStack` (fused kernels), owned by VisionTransformer / CoCaTextDecoder /
CoCaMultimodalDecoder; `TransformerEncoderLayer` / `TransformerEncoder` are also callable on their own (forward values,
same kernels: `engine_layers.py`).  The decoder classes stay containers (KV-cache decoding is outside the path)."""
from typing import Any, Callable, List, NamedTuple, Optional, Tuple

from torch import nn, Tensor

from ..._lib import MMBError


class TransformerOutput(NamedTuple):
    last_hidden_state: Optional[Tensor] = None
    pooler_output: Optional[Tensor] = None
    hidden_states: Optional[List[Tensor]] = None
    attentions: Optional[List[Tensor]] = None
    image_labels: Optional[Tensor] = None
    current_key_values: Optional[List[Tuple[Tensor, Tensor]]] = None


def _no_dropout(dropout: float, what: str) -> None:
    if dropout != 0.0:
        raise NotImplementedError(f"{what}: dropout > 0 is not on the accelerated path (reference default is 0.0)")


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model: int, n_head: int, dim_feedforward: int, dropout: float = 0.0,
                 activation: Callable[..., nn.Module] = nn.ReLU, layer_norm_eps: float = 1e-12, norm_first: bool = False,
                 drop_path_rate: Optional[float] = None) -> None:
        super().__init__()
        from .mlp import MLP
        from .multi_head_attention import MultiHeadSelfAttention
        from .normalizations import Fp32LayerNorm

        _no_dropout(dropout, "TransformerEncoderLayer")
        if drop_path_rate is not None:
            raise NotImplementedError("stochastic depth (drop_path_rate) is not on the accelerated path")
        self.attention = MultiHeadSelfAttention(embed_dim=d_model, num_heads=n_head)
        self.attention_dropout = nn.Dropout(dropout)
        self.feedforward_dropout = nn.Dropout(dropout)
        self.feedforward = MLP(d_model, d_model, dim_feedforward, dropout=dropout, activation=activation)
        self.attention_layernorm = Fp32LayerNorm(d_model, eps=layer_norm_eps)
        self.feedforward_layernorm = Fp32LayerNorm(d_model, eps=layer_norm_eps)
        self.norm_first = norm_first

    def forward(self, hidden_states: Tensor, attention_mask: Optional[Tensor] = None) -> Tensor:
        """Standalone forward (values only; inside VisionTransformer the layer runs in the fused LayerStack)."""
        from ...engine_layers import encoder_layer_forward

        return encoder_layer_forward(self, hidden_states, attention_mask)


class TransformerEncoder(nn.Module):
    def __init__(self, n_layer: int, d_model: int, n_head: int, dim_feedforward: int, dropout: float = 0.0,
                 activation: Callable[..., nn.Module] = nn.ReLU, layer_norm_eps: float = 1e-12, norm_first: bool = False,
                 final_layer_norm_eps: Optional[float] = None, drop_path_rate: Optional[float] = None):
        super().__init__()
        from .normalizations import Fp32LayerNorm

        self.layer = nn.ModuleList([
            TransformerEncoderLayer(d_model, n_head, dim_feedforward, dropout, activation, layer_norm_eps, norm_first,
                                    drop_path_rate) for _ in range(n_layer)])
        self.final_layer_norm = None
        if final_layer_norm_eps:
            self.final_layer_norm = Fp32LayerNorm(d_model, eps=final_layer_norm_eps)

    def forward(self, hidden_states: Tensor, attention_mask: Optional[Tensor] = None,
                return_hidden_states: bool = False) -> TransformerOutput:
        """Standalone forward (values only; inside VisionTransformer the stack runs in the fused LayerStack)."""
        from ...engine_layers import encoder_forward

        return encoder_forward(self, hidden_states, attention_mask, return_hidden_states)


class TransformerDecoderLayer(nn.Module):
    def __init__(self, d_model: int, n_head: int, dim_feedforward: int, dropout: float = 0.0,
                 activation: Callable[..., nn.Module] = nn.ReLU, layer_norm_eps: float = 1e-12, norm_first: bool = False,
                 use_cross_attention: bool = True, dim_kv: Optional[int] = None) -> None:
        super().__init__()
        from .mlp import MLP
        from .multi_head_attention import MultiHeadAttentionWithCache
        from .normalizations import Fp32LayerNorm

        _no_dropout(dropout, "TransformerDecoderLayer")
        dim_kv = dim_kv if dim_kv is not None else d_model
        self.attention = MultiHeadAttentionWithCache(dim_q=d_model, dim_kv=d_model, num_heads=n_head, dropout=dropout)
        self.attention_dropout = nn.Dropout(dropout)
        self.cross_attention: Optional[MultiHeadAttentionWithCache] = None
        self.use_cross_attention = use_cross_attention
        if self.use_cross_attention:
            self.cross_attention = MultiHeadAttentionWithCache(dim_q=d_model, dim_kv=dim_kv, num_heads=n_head,
                                                               dropout=dropout)
            self.cross_attention_layernorm = Fp32LayerNorm(d_model, eps=layer_norm_eps)
            self.cross_attention_dropout = nn.Dropout(dropout)
        self.feedforward = MLP(d_model, d_model, dim_feedforward, dropout=dropout, activation=activation)
        self.feedforward_dropout = nn.Dropout(dropout)
        self.attention_layernorm = Fp32LayerNorm(d_model, eps=layer_norm_eps)
        self.feedforward_layernorm = Fp32LayerNorm(d_model, eps=layer_norm_eps)
        self.norm_first = norm_first

    def forward(self, *args: Any, **kwargs: Any) -> Tuple[Tensor, Optional[Tuple[Tensor, Tensor]]]:
        raise MMBError("TransformerDecoderLayer runs inside its decoder's fused runtime; not a standalone op here")


class TransformerDecoder(nn.Module):
    def __init__(self, n_layer: int, d_model: int, n_head: int, dim_feedforward: int, dropout: float = 0.0,
                 activation: Callable[..., nn.Module] = nn.ReLU, layer_norm_eps: float = 1e-12, norm_first: bool = False,
                 use_cross_attention: bool = True, dim_kv: Optional[int] = None,
                 final_layer_norm_eps: Optional[float] = None, cross_attention_interval: int = 1):
        super().__init__()
        from .normalizations import Fp32LayerNorm

        if use_cross_attention and cross_attention_interval != 1:
            raise NotImplementedError("cross_attention_interval != 1 is not on the accelerated path")
        self.layer = nn.ModuleList([
            TransformerDecoderLayer(d_model, n_head, dim_feedforward, dropout, activation, layer_norm_eps, norm_first,
                                    use_cross_attention and (i % cross_attention_interval == 0), dim_kv)
            for i in range(n_layer)])
        self.final_layer_norm = None
        if final_layer_norm_eps:
            self.final_layer_norm = Fp32LayerNorm(d_model, eps=final_layer_norm_eps)

    def forward(self, *args: Any, **kwargs: Any) -> TransformerOutput:
        raise MMBError("TransformerDecoder runs inside its owner's fused runtime; not a standalone op here")
