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
