"""CPU restatement (plain fp32 torch tensor math on a state dict) of the FLAVA pre-training losses — TEST
INFRASTRUCTURE ONLY: imported by tests/ (parity checker) and never by the product path.

Follows torchmultimodal/modules/losses/flava.py: Pooler :84-97, TwoWayHead :100-107, ITMLoss :110-140,
MaskedPredictionHead :143-179, MaskedPredictionLoss :182-238, FLAVAGlobalContrastiveLoss :241-293,
FLAVAPretrainingLoss.forward :370-484.  Pinned against the reference's own outputs in
tests/golden/flava_pretraining_golden.pt (tests/test_flava_pretraining_cpu.py).
"""
import math
from typing import Dict, Optional

import torch
from torch import Tensor

from . import clip_oracle as O


def _lin(x: Tensor, sd: Dict[str, Tensor], p: str, bias_key: Optional[str] = None) -> Tensor:
    y = x @ sd[p + ".weight"].t()
    b = sd.get(bias_key if bias_key is not None else p + ".bias")
    return y + b if b is not None else y


def gelu(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))      # nn.functional.gelu (erf form)


def ce_ignore(logits: Tensor, labels: Tensor, ignore_index: int = -1) -> Tensor:
    """nn.CrossEntropyLoss(ignore_index): mean over the rows whose label is kept (0/0 = NaN when none is)."""
    keep = labels != ignore_index
    lsm = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    nll = -lsm.gather(1, labels.clamp_min(0).view(-1, 1)).squeeze(1)
    return (nll * keep).sum() / keep.sum()


def masked_prediction(hidden: Tensor, labels: Optional[Tensor], sd: Dict[str, Tensor], p: str, eps: float = 1e-5,
                      ignore_index: int = -1):
    """MaskedPredictionLoss.forward (:206-238) with MaskedPredictionHead (:174-179)."""
    if labels is not None:
        keep = labels.ne(ignore_index)                            # :212-215
        labels = labels[keep]
        x = hidden[keep, :]
    else:
        x = hidden
    x = gelu(_lin(x, sd, p + ".cls.dense"))                       # :175-176
    x = O.layer_norm(x, sd[p + ".cls.layer_norm.weight"], sd[p + ".cls.layer_norm.bias"], eps)   # :177
    logits = x @ sd[p + ".cls.decoder.weight"].t() + sd[p + ".cls.bias"]                          # :178, :164-172
    if labels is None:
        return logits, logits.sum() * 0
    return logits, ce_ignore(logits.view(-1, logits.shape[-1]), labels.view(-1), ignore_index)


def itm(hidden: Tensor, labels: Optional[Tensor], sd: Dict[str, Tensor], p: str = "itm_loss", ignore_index: int = -1):
    pooled = torch.tanh(_lin(hidden[:, 0], sd, p + ".pooler.dense"))          # :91-97
    scores = _lin(pooled, sd, p + ".cls.seq_relationship")                     # :106-107
    if labels is None:
        return scores, pooled.sum() * 0
    return scores, ce_ignore(scores.view(-1, 2), labels.view(-1), ignore_index)


def global_contrastive(image: Tensor, text: Tensor, mask: Optional[Tensor], sd: Dict[str, Tensor],
                       p: str = "contrastive_loss"):
    t = O.normalize(text)                                                      # :267-271 (dim=-1)
    i = O.normalize(image)
    scale = sd[p + ".logit_scale"].clamp(0, 4.6052)                            # :273
    loss, la, lb, loss_a, loss_b = O.contrastive_loss(i, t, scale, mask=mask)  # :275-282
    return dict(loss=loss, image_logits=la, text_logits=lb, image_loss=loss_a, text_loss=loss_b, image_embedding=i,
                text_embedding=t)


def pretraining_loss(sd: Dict[str, Tensor], *, image_sequence=None, text_sequence=None, image_masked_sequence=None,
                     text_masked_sequence=None, multimodal_sequence=None, multimodal_masked_sequence=None, itm_labels=None,
                     mim_labels=None, mlm_labels=None, projected_image_embeddings=None, projected_text_embeddings=None,
                     weights: Optional[Dict[str, float]] = None) -> Dict[str, Tensor]:
    """FLAVAPretrainingLoss.forward (:370-484); returns the flat dict of tests/flava_pretraining_cases.flatten_loss_output."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    w = dict(mlm=1.0, mim=1.0, contrastive=1.0, mmm_image=1.0, mmm_text=1.0, itm=1.0)
    w.update(weights or {})
    out: Dict[str, Tensor] = {}
    pos_mask = None
    if image_masked_sequence is not None and w["mim"] > 0 and multimodal_masked_sequence is None:      # :386-399
        start = -mim_labels.size(1) if mim_labels is not None else 1
        lg, ls = masked_prediction(image_masked_sequence[:, start:, :], mim_labels, sd, "mim_loss")
        out["mim_output.logits"], out["mim_output.loss"] = lg, ls * w["mim"]
        out["losses.mim_loss"] = out["mim_output.loss"]
    if text_masked_sequence is not None and w["mlm"] > 0 and multimodal_masked_sequence is None:       # :402-413
        start = -mlm_labels.size(1) if mlm_labels is not None else 1
        lg, ls = masked_prediction(text_masked_sequence[:, start:, :], mlm_labels, sd, "mlm_loss")
        out["mlm_output.logits"], out["mlm_output.loss"] = lg, ls * w["mlm"]
        out["losses.mlm_loss"] = out["mlm_output.loss"]
    if multimodal_masked_sequence is not None and w["itm"] > 0:                                         # :415-435
        if itm_labels is not None:
            pos_pairs = itm_labels.ne(0)
            pos_mask = torch.where(pos_pairs.any(), pos_pairs, pos_pairs.new([True]))
        else:
            pos_mask = torch.ones(multimodal_masked_sequence.size(0)).bool()
        sc, ls = itm(multimodal_masked_sequence, itm_labels, sd)
        out["itm_output.logits"], out["itm_output.loss"] = sc, ls * w["itm"]
        out["losses.itm_loss"] = out["itm_output.loss"]
        multimodal_masked_sequence = multimodal_masked_sequence[pos_mask]
        if mlm_labels is not None:
            mlm_labels = mlm_labels[pos_mask]
        if mim_labels is not None:
            mim_labels = mim_labels[pos_mask]
    if multimodal_masked_sequence is not None and w["mmm_text"] > 0:                                    # :437-449
        start = -mlm_labels.size(1) if mlm_labels is not None else -(text_masked_sequence.size(1) - 1)
        lg, ls = masked_prediction(multimodal_masked_sequence[:, start:, :], mlm_labels, sd, "mmm_loss.mlm")
        out["mmm_text_output.logits"], out["mmm_text_output.loss"] = lg, ls * w["mmm_text"]
        out["losses.mmm_text_loss"] = out["mmm_text_output.loss"]
    if multimodal_masked_sequence is not None and w["mmm_image"] > 0:                                   # :451-466
        total = mim_labels.size(1) if mlm_labels is not None else (image_masked_sequence.size(1) - 1)
        lg, ls = masked_prediction(multimodal_masked_sequence[:, 2:2 + total, :], mim_labels, sd, "mmm_loss.mim")
        out["mmm_image_output.logits"], out["mmm_image_output.loss"] = lg, ls * w["mmm_image"]
        out["losses.mmm_image_loss"] = out["mmm_image_output.loss"]
    if projected_image_embeddings is not None and projected_text_embeddings is not None and w["contrastive"] > 0:   # :468-482
        gc = global_contrastive(projected_image_embeddings, projected_text_embeddings, pos_mask, sd)
        gc["loss"] = gc["loss"] * w["contrastive"]
        for k, v in gc.items():
            out[f"global_contrastive_output.{k}"] = v
        out["losses.global_contrastive_loss"] = gc["loss"]
    return out
