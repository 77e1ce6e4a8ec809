"""GPU parity of the FLAVA pre-training losses (reference modules/losses/flava.py:84-484,
models/flava/model.py:300-377) against outputs of the unmodified reference (tests/golden/flava_pretraining_golden.pt)
and the oracle.

Tolerance: bf16 GEMM operands with fp32 accumulation, fp32 LayerNorm / softmax statistics: logits agree to 2e-2 of their
absmax, losses (values 0.6 - 7.3) to 2e-2 absolute, unit-norm embeddings to 5e-3.  Row selection, label handling and the
positive-pair mask are exact (shapes and row order are asserted)."""
import os

import pytest
import torch

import flava_pretraining_cases as PC
from oracle import flava_loss_oracle as LO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "flava_pretraining_golden.pt")


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLD)


def _compare(out, ref):
    assert set(out) == set(ref), sorted(set(out) ^ set(ref))
    for k, r in ref.items():
        o = out[k]
        assert o.shape == r.shape, (k, o.shape, r.shape)
        if r.dim() == 0:
            assert abs(o.item() - r.item()) < 2e-2, (k, o.item(), r.item())
        elif k.endswith("_embedding"):
            assert (o - r).abs().max().item() < 5e-3, k
        else:
            err = (o - r).abs().max().item() / max(r.abs().max().item(), 1e-6)
            assert err < 2e-2, (k, err)


def _to(x, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in x.items()}


@pytest.mark.parametrize("name", ["unimodal", "multimodal"])
def test_pretraining_loss_against_reference_golden(dev, golden, name):
    from multimodal_b200._lib import MMBError
    from multimodal_b200.modules.losses.flava import FLAVAPretrainingLoss

    m = PC.build_loss(FLAVAPretrainingLoss).to(dev)
    kw = _to(PC.loss_calls()[name], dev)
    with torch.no_grad():
        o = m(**kw)
    _compare(PC.flatten_loss_output(o), golden[f"loss.{name}"])
    o2 = m(**kw)                           # grad mode on + trainable heads: the autograd path, same values
    assert all(v.requires_grad for v in o2.losses.values() if v is not None)
    _compare(PC.flatten_loss_output(o2), golden[f"loss.{name}"])


def test_flava_for_pretraining_against_reference_golden(dev, golden):
    from multimodal_b200.models.flava import flava_model, FLAVAForPreTraining
    from multimodal_b200.modules.losses.flava import FLAVAPretrainingLoss

    m = PC.build_model(flava_model, FLAVAForPreTraining, FLAVAPretrainingLoss)
    assert PC.param_checksum(m) == pytest.approx(golden["model_param_checksum"], rel=1e-12)
    m = m.to(dev)
    inp, _ = PC.model_inputs()
    with torch.no_grad():
        o = m(**_to(inp, dev))
    _compare(PC.flatten_loss_output(o), golden["model"])


def test_masked_prediction_loss_real_shapes_against_oracle(dev):
    """FLAVA's real head shapes: hidden 768, text vocabulary 30 522 (not a multiple of 8), ~15 % of 8 x 128 tokens kept."""
    from multimodal_b200.modules.losses.flava import MaskedPredictionLoss

    torch.manual_seed(3)
    m = MaskedPredictionLoss(hidden_size=768, vocab_size=30522).eval()
    with torch.no_grad():
        m.cls.bias.add_(0.1 * torch.randn(30522))
        m.cls.layer_norm.weight.add_(0.1 * torch.randn(768))
    g = torch.Generator().manual_seed(4)
    hidden = torch.randn(8, 129, 768, generator=g)[:, 1:, :]            # a slice view, as the pre-training loss passes it
    labels = torch.full((8, 128), -1, dtype=torch.long)
    pick = torch.rand(8, 128, generator=g) < 0.15
    labels[pick] = torch.randint(0, 30522, (int(pick.sum()),), generator=g)
    sd = {"x." + k: v for k, v in m.state_dict().items()}
    ref_logits, ref_loss = LO.masked_prediction(hidden, labels, {k: v.float() for k, v in sd.items()}, "x")
    m = m.to(dev)
    with torch.no_grad():
        o = m(hidden.to(dev), labels.to(dev))
    assert o.logits.shape == ref_logits.shape
    err = (o.logits.cpu() - ref_logits).abs().max().item() / ref_logits.abs().max().item()
    assert err < 2e-2, err
    assert abs(o.loss.item() - ref_loss.item()) < 2e-2
    # no label kept: CrossEntropyLoss over zero rows is NaN in the reference too; ignore_nan turns it into 0
    with torch.no_grad():
        assert torch.isnan(m(hidden.to(dev), torch.full((8, 128), -1, dtype=torch.long, device=dev)).loss)
