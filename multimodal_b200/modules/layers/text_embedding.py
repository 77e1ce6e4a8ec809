"""Parameter container mirroring torchmultimodal/modules/layers/text_embedding.py:13-104 (BERTTextEmbeddings).
The sum of the three gathers + LayerNorm is one kernel (`mmb_bert_embed_ln_fwd`) inside BERTTextEncoder."""
from torch import nn, Tensor

from ..._lib import MMBError


class BERTTextEmbeddings(nn.Module):
    def __init__(self, hidden_size: int = 768, vocab_size: int = 30522, pad_token_id: int = 0,
                 max_position_embeddings: int = 512, type_vocab_size: int = 2, layer_norm_eps: float = 1e-12,
                 dropout: float = 0.0, offset_pos_ids: bool = False) -> None:
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab_size, hidden_size, pad_token_id)
        self.position_embeddings = nn.Embedding(max_position_embeddings, hidden_size)
        self.token_type_embeddings = nn.Embedding(type_vocab_size, hidden_size)
        self.layer_norm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.dropout = nn.Dropout(dropout)
        self.pad_token_id = pad_token_id
        self.offset_pos_ids = offset_pos_ids
        if offset_pos_ids:
            raise NotImplementedError("RoBERTa-style offset position ids are not on the FLAVA path")

    def forward(self, *args, **kwargs) -> Tensor:
        raise MMBError("BERTTextEmbeddings is fused into BERTTextEncoder's runtime; not a standalone op here")
