"""`flava_text_encoder` — same builder signature as torchmultimodal/models/flava/text_encoder.py:22-71."""
from functools import partial
from typing import Callable

from torch import nn

from ...modules.encoders.bert_text_encoder import BERTTextEncoder
from ...modules.layers.normalizations import Fp32LayerNorm
from ...modules.layers.text_embedding import BERTTextEmbeddings
from ...modules.losses.flava import Pooler
from .transformer import init_transformer_weights, TransformerEncoder


def flava_text_encoder(num_hidden_layers: int = 12, hidden_size: int = 768, num_attention_heads: int = 12,
                       intermediate_size: int = 3072, intermediate_activation: Callable[..., nn.Module] = nn.GELU,
                       layer_norm_eps: float = 1e-12, dropout: float = 0.0, vocab_size: int = 30522,
                       pad_token_id: int = 0, type_vocab_size: int = 2, max_position_embeddings: int = 512,
                       initializer_range: float = 0.02) -> BERTTextEncoder:
    # Sub-modules are built in the reference's order (embeddings, encoder, final LayerNorm, pooler): keyword values are
    # evaluated left to right, so a seeded build consumes the RNG exactly like the reference builder does.
    common = dict(layer_norm_eps=layer_norm_eps, dropout=dropout)
    return BERTTextEncoder(
        embeddings=BERTTextEmbeddings(hidden_size=hidden_size, vocab_size=vocab_size, pad_token_id=pad_token_id,
                                      type_vocab_size=type_vocab_size, max_position_embeddings=max_position_embeddings,
                                      **common),
        encoder=TransformerEncoder(n_layer=num_hidden_layers, d_model=hidden_size, n_head=num_attention_heads,
                                   dim_feedforward=intermediate_size, activation=intermediate_activation, norm_first=True,
                                   **common),
        layernorm=Fp32LayerNorm(hidden_size, eps=layer_norm_eps),
        pooler=Pooler(hidden_size=hidden_size),
        weight_init_fn=partial(init_transformer_weights, initializer_range=initializer_range))
