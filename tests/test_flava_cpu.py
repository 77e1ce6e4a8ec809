"""CPU tests of the FLAVA encoder path: the oracle pinned to the reference goldens, and the drop-in parameter
containers (state-dict keys / seeded init) pinned through the parameter checksum the fixture generator recorded from
the REFERENCE model."""
import os

import pytest
import torch

import flava_cases as FC
from oracle import flava_oracle as FO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "flava_golden.pt")


@pytest.fixture(scope="module")
def flava_golden():
    return torch.load(GOLD)


@pytest.mark.parametrize("name", list(FC.CASES))
def test_flava_oracle_and_init_mirror_match_reference_golden(flava_golden, name):
    from multimodal_b200.models.flava import flava_model

    g = flava_golden[name]
    m = FC.build(flava_model, name)
    # same parameter set, same creation order, same RNG consumption as the reference builder
    assert FC.param_checksum(m) == pytest.approx(g["param_checksum"], rel=1e-12)
    inp = FC.inputs(name)
    for k, v in g["inputs"].items():
        assert torch.equal(inp[k], v)
    out = FO.flava_forward(m.state_dict(), FC.CASES[name]["kwargs"], inp["image"], inp["text"],
                           inp["image_patches_mask"], inp["text_masked"], skip_unmasked_mm_encoder=False)
    assert set(out) == set(g["outputs"])
    for k, ref in g["outputs"].items():
        assert torch.allclose(out[k], ref, rtol=1e-5, atol=1e-5), k


def test_flava_state_dict_keys_follow_reference_layout():
    from multimodal_b200.models.flava import flava_model

    m = FC.build(flava_model, "flava_small")
    keys = set(m.state_dict())
    for k in ("image_encoder.embeddings.cls_token", "image_encoder.embeddings.mask_token",
              "image_encoder.embeddings.patch_embeddings.projection.weight",
              "image_encoder.encoder.layer.1.attention.query.weight", "image_encoder.encoder.layer.0.attention.output.bias",
              "image_encoder.encoder.layer.0.feedforward.model.0.weight", "image_encoder.encoder.layer.0.feedforward.model.2.bias",
              "image_encoder.encoder.layer.0.attention_layernorm.weight", "image_encoder.encoder.layer.0.feedforward_layernorm.bias",
              "image_encoder.layernorm.weight", "image_encoder.pooler.dense.weight",
              "text_encoder.embeddings.word_embeddings.weight", "text_encoder.embeddings.position_embeddings.weight",
              "text_encoder.embeddings.token_type_embeddings.weight", "text_encoder.embeddings.layer_norm.weight",
              "mm_encoder.cls_token", "mm_encoder.encoder.layer.0.attention.key.weight",
              "image_to_mm_projection.weight", "text_to_mm_projection.bias", "text_projection.weight", "image_projection.bias"):
        assert k in keys, k


def test_flava_modules_refuse_cpu_execution():
    """No CPU fallback: the encoders fail loudly when asked to run off-GPU."""
    from multimodal_b200._lib import MMBError
    from multimodal_b200.models.flava import flava_model

    m = FC.build(flava_model, "flava_small")
    inp = FC.inputs("flava_small")
    with pytest.raises(MMBError):
        m(image=inp["image"], text=inp["text"])
    with pytest.raises(ValueError):
        m.image_encoder(None)
    with pytest.raises(ValueError):
        m.text_encoder()
