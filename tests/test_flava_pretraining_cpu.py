"""CPU tests of the FLAVA pre-training losses: the oracle restatement pinned to outputs of the UNMODIFIED
reference classes (tests/golden/flava_pretraining_golden.pt), and the drop-in parameter containers pinned through the
parameter checksum the fixture generator recorded from the reference modules."""
import os

import pytest
import torch

import flava_pretraining_cases as PC
from oracle import flava_loss_oracle as LO
from oracle import flava_oracle as FO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "flava_pretraining_golden.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLD)


def _close(out, ref, rtol=1e-5, atol=2e-5):
    assert set(out) == set(ref), (sorted(set(out) ^ set(ref)))
    for k, r in ref.items():
        assert out[k].shape == r.shape, k
        assert torch.allclose(out[k], r, rtol=rtol, atol=atol), (k, (out[k] - r).abs().max())


@pytest.mark.parametrize("name", ["unimodal", "multimodal"])
def test_loss_oracle_and_init_mirror_match_reference_golden(golden, name):
    from multimodal_b200.modules.losses.flava import FLAVAPretrainingLoss

    m = PC.build_loss(FLAVAPretrainingLoss)
    # same parameter set, same creation order, same RNG consumption as the reference module
    assert PC.param_checksum(m) == pytest.approx(golden["loss_param_checksum"], rel=1e-12)
    out = LO.pretraining_loss(m.state_dict(), **PC.loss_calls()[name])
    _close(out, golden[f"loss.{name}"])


def test_loss_state_dict_keys_follow_reference_layout():
    from multimodal_b200.modules.losses.flava import FLAVAPretrainingLoss

    keys = set(FLAVAPretrainingLoss(hidden_size=128, text_vocab_size=40, image_vocab_size=24).state_dict())
    for k in ("contrastive_loss.logit_scale", "mlm_loss.cls.dense.weight", "mlm_loss.cls.layer_norm.bias",
              "mlm_loss.cls.decoder.weight", "mlm_loss.cls.bias", "mim_loss.cls.decoder.weight",
              "mmm_loss.mlm.cls.dense.bias", "mmm_loss.mim.cls.bias", "itm_loss.pooler.dense.weight",
              "itm_loss.cls.seq_relationship.bias"):
        assert k in keys, k
    # decoder.bias aliases the output-only bias (losses/flava.py:164-172): one tensor, two names
    assert "mlm_loss.cls.decoder.bias" in keys


def test_whole_model_oracle_matches_reference_golden(golden):
    from multimodal_b200.models.flava import flava_model, FLAVAForPreTraining
    from multimodal_b200.modules.losses.flava import FLAVAPretrainingLoss

    m = PC.build_model(flava_model, FLAVAForPreTraining, FLAVAPretrainingLoss)
    assert PC.param_checksum(m) == pytest.approx(golden["model_param_checksum"], rel=1e-12)
    inp, _ = PC.model_inputs()
    labels = m.image_codebook(inp["image_for_codebook"]).flatten(1)
    pm = inp["image_patches_mask"].flatten(1).to(torch.bool)
    labels[~pm] = -1
    sd = m.state_dict()
    f = FO.flava_forward({k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}, PC.MODEL_KW, inp["image"],
                         inp["text"], pm, inp["text_masked"], skip_unmasked_mm_encoder=False)
    out = LO.pretraining_loss(
        {k[len("loss."):]: v for k, v in sd.items() if k.startswith("loss.")},
        image_sequence=f["image.last_hidden_state"], text_sequence=f["text.last_hidden_state"],
        image_masked_sequence=f["image_masked.last_hidden_state"], text_masked_sequence=f["text_masked.last_hidden_state"],
        multimodal_sequence=f["multimodal.last_hidden_state"],
        multimodal_masked_sequence=f["multimodal_masked.last_hidden_state"], itm_labels=inp["itm_labels"],
        mim_labels=labels, mlm_labels=inp["mlm_labels"], projected_image_embeddings=f["projected_image_embeddings"],
        projected_text_embeddings=f["projected_text_embeddings"])
    _close(out, golden["model"], rtol=1e-4, atol=1e-4)
