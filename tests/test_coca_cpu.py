"""CPU tests of the CoCa forward path: oracle pinned to the reference goldens; drop-in containers pinned through the
parameter checksum recorded from the REFERENCE model (same keys, creation order, RNG consumption)."""
import os

import pytest
import torch

import coca_cases as CC
from oracle import coca_oracle as CO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "coca_golden.pt")


@pytest.mark.parametrize("name", list(CC.CASES))
def test_coca_oracle_and_init_mirror_match_reference_golden(name):
    from multimodal_b200.models.coca import coca_for_pretraining

    g = torch.load(GOLD)[name]
    m = CC.build(coca_for_pretraining, name)
    assert CC.param_checksum(m) == pytest.approx(g["param_checksum"], rel=1e-12)
    inp = CC.inputs(name)
    for k, v in g["inputs"].items():
        assert torch.equal(inp[k], v)
    out = CO.coca_forward(m.state_dict(), CC.CASES[name]["kwargs"], inp["images"], inp["texts"])
    for k in ("image_pooled_output", "text_pooled_output", "multimodal_embeddings"):
        assert out[k].shape == g[k].shape, k
        assert torch.allclose(out[k], g[k], rtol=1e-4, atol=2e-5), (k, (out[k] - g[k]).abs().max())
    if "contrastive" in g:
        assert torch.allclose(out["contrastive"], g["contrastive"], atol=1e-4)
        assert torch.allclose(out["captioning"], g["captioning"], atol=1e-4)


def test_coca_text_mask_matches_reference_semantics():
    """build_mask of the drop-in module == the oracle's restatement == hand-built expectation (text_decoder.py:141-162)."""
    from multimodal_b200.models.coca.text_decoder import CoCaTextDecoder

    dec = CoCaTextDecoder(vocab_size=50, num_positions=5, embedding_dim=128, n_layer=1, n_head=2, dim_feedforward=128,
                          output_dim=128)
    ids = torch.tensor([[3, 4, 0, 0], [5, 6, 7, 8]])
    m = dec.build_mask(ids)
    assert m.shape == (2, 1, 5, 5)
    assert torch.equal(m.bool(), CO.text_mask(ids, 5))
    # text rows: pure causal; CLS row (last): column 0 always, column j+1 iff token j is not padding
    assert m[0, 0, :4].bool().tolist() == torch.tril(torch.ones(5, 5))[:4].bool().tolist()
    assert m[0, 0, 4].bool().tolist() == [True, True, True, False, False]
    assert m[1, 0, 4].bool().tolist() == [True] * 5


def test_coca_modules_refuse_cpu_execution():
    from multimodal_b200._lib import MMBError
    from multimodal_b200.models.coca import coca_for_pretraining

    m = CC.build(coca_for_pretraining, "coca_parallel")
    inp = CC.inputs("coca_parallel")
    with pytest.raises(MMBError):
        m(inp["images"], inp["texts"])


def test_coca_oracle_reproduces_reference_known_answers():
    """The reference's own constant-init known-answer test (tests/models/coca/test_coca_model.py:45-165): pooled
    embeddings 0.3536, logits 8.0, losses contrastive 0.6931 / captioning 3.9120 (parallel pooler), run through OUR
    drop-in builder's state dict (same keys) and the oracle."""
    from multimodal_b200.models.coca import coca_for_pretraining

    kw = dict(vision_patch_size=4, vision_dim_feedforward=24, vision_n_layer=2, vision_n_head=2, vocab_size=50,
              num_text_positions=11, text_hidden_dim=8, text_n_layer=2, text_n_head=2, text_dim_feedforward=32,
              text_output_dim=8, fusion_n_layer=2, fusion_n_head=2, fusion_dim_feedforward=32,
              multimodal_output_projection_dim=50, pooler_input_embed_dim=6, pooler_output_embed_dim=8, image_size=12,
              pooler_n_head=2, cascaded_pooler=False)
    m = coca_for_pretraining(**kw).eval()
    with torch.no_grad():   # tests/test_utils.py:193-205 init_weights_with_constant
        for n, p in m.named_parameters():
            p.fill_(0.0 if n.endswith(("text_projection.bias", "output_projection.bias", "vision_proj.bias")) else 1.0)
    torch.manual_seed(0)
    images = torch.randn(2, 3, 12, 12)
    texts = torch.tensor([[1, 3, 4, 5, 6, 7, 8, 2, 0, 0, 0], [1, 25, 28, 34, 39, 45, 40, 5, 12, 6, 2]])
    out = CO.coca_forward(m.state_dict(), kw, images, texts)
    assert torch.allclose(out["image_pooled_output"], 0.3536 * torch.ones(2, 8), atol=1e-4)
    assert torch.allclose(out["text_pooled_output"], 0.3536 * torch.ones(2, 8), atol=1e-4)
    assert torch.allclose(out["multimodal_embeddings"], 8.0 * torch.ones(2, 10, 50), atol=1e-4)
    assert abs(out["contrastive"].item() - 0.6931) < 1e-4
    assert abs(out["captioning"].item() - 3.9120) < 1e-4
