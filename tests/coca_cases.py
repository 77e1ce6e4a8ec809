"""Shared definition of the CoCa parity case (see tests/flava_cases.py for the scheme)."""
import torch

CASES = {
    # vision: 64 tokens, d=256 (4 heads of 64); poolers d_out=384 with 4 heads -> head_dim 96 (as CoCa ViT-L/14);
    # text decoder 13 positions (12 tokens + CLS), d=384 / 6 heads; fusion 6 heads + cross-attention to 32 queries
    "coca_small": dict(
        kwargs=dict(vision_patch_size=4, vision_dim_feedforward=512, vision_n_layer=2, vision_n_head=4, image_size=32,
                    vocab_size=512, num_text_positions=13, text_hidden_dim=384, text_n_layer=2, text_n_head=6,
                    text_dim_feedforward=768, text_output_dim=384, fusion_n_layer=2, fusion_n_head=6,
                    fusion_dim_feedforward=768, multimodal_output_projection_dim=512, pooler_input_embed_dim=256,
                    pooler_output_embed_dim=384, pooler_n_head=4, pooler_n_queries=32, cascaded_pooler=True),
        batch=4),
    # parallel pooler (n_queries + 1 queries, query 0 = contrastive): the configuration the reference's own
    # CoCaForPretraining test uses (tests/models/coca/test_coca_model.py:55-122); losses are pinned on this one
    "coca_parallel": dict(
        kwargs=dict(vision_patch_size=8, vision_dim_feedforward=256, vision_n_layer=1, vision_n_head=2, image_size=32,
                    vocab_size=300, num_text_positions=9, text_hidden_dim=128, text_n_layer=1, text_n_head=2,
                    text_dim_feedforward=256, text_output_dim=128, fusion_n_layer=1, fusion_n_head=2,
                    fusion_dim_feedforward=256, multimodal_output_projection_dim=300, pooler_input_embed_dim=128,
                    pooler_output_embed_dim=128, pooler_n_head=2, pooler_n_queries=16, cascaded_pooler=False),
        batch=8),
}


def build(builder, name: str):
    torch.manual_seed(0)
    m = builder(**CASES[name]["kwargs"])
    g = torch.Generator().manual_seed(13)
    with torch.no_grad():
        for n, p in m.named_parameters():
            # keep attention logits O(1): the reference init of text_projection has std d**0.5 (huge), leave it alone
            p.add_(0.03 * torch.randn(p.shape, generator=g))
    return m.eval()


def param_checksum(m) -> float:
    return float(sum(p.detach().double().abs().sum() for p in m.parameters()))


def inputs(name: str):
    c = CASES[name]
    kw, B = c["kwargs"], c["batch"]
    g = torch.Generator().manual_seed(6)
    images = torch.randn(B, 3, kw["image_size"], kw["image_size"], generator=g)
    T = kw["num_text_positions"]
    texts = torch.randint(1, kw["vocab_size"], (B, T), generator=g)
    for b in range(B):  # ragged right padding with pad_idx = 0
        n_pad = (4 * b) % (T - 3)
        if n_pad:
            texts[b, T - n_pad:] = 0
    return dict(images=images, texts=texts)
