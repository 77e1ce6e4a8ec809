"""FLAVA pre-training losses — drop-in for torchmultimodal/modules/losses/flava.py:22-484.

Same classes, constructor arguments, state-dict keys and initialisation order (`Pooler`, `TwoWayHead`, `ITMLoss`,
`MaskedPredictionHead`, `MaskedPredictionLoss`, `FLAVAGlobalContrastiveLoss`, `FLAVAPretrainingLoss`) and the same output
containers.  The modules are parameter containers; the arithmetic runs on the library's kernels:

* masked-prediction heads (MLM / MIM / MMM, :143-238): boolean-mask row select (`mmb_gather_rows_idx_cast`) ->
  dense GEMM with bias + erf-GELU epilogue -> fp32 LayerNorm -> vocabulary GEMM + bias (fp32 logits, returned as in the
  reference) -> `nn.CrossEntropyLoss(ignore_index)` (`mmb_ce_labels`);
* ITM head (:84-140): first-token select -> dense GEMM + bias -> tanh -> Linear(hidden, 2) -> cross-entropy;
* global contrastive loss (:241-293): `F.normalize` kernels + the fused similarity-GEMM / cross-entropy runtime of
  `contrastive_loss_with_temperature` with the positive-pair mask.

With grad mode on and trainable parameters (or inputs that require grad) the same modules build an autograd graph:
the heads run as the Functions of ``engine_flava_heads`` (one fused node per masked-prediction head, explicit backward
schedules on the library's kernels), the pooler as ``engine_flava_train.FirstTokenLinearFunction``, the contrastive loss
through ``engine_loss.ContrastiveFunction``.  Under ``torch.no_grad()`` they compute forward values only.
GEMM operands are bf16 (fp32 accumulate, fp32 LayerNorm / softmax statistics); tolerances in
tests/test_gpu_flava_pretraining.py.
"""
import math
import warnings
from collections import OrderedDict
from dataclasses import field, fields, make_dataclass
from typing import Any, Callable, Optional, Union

import torch
from torch import nn, Tensor

from ... import ops
from ..._lib import MMBError
from ...utils.distributed import BackpropType
from ..layers.normalizations import Fp32LayerNorm
from .contrastive_loss_with_temperature import contrastive_loss_with_temperature


def assert_labels_are_present(labels: Optional[Tensor], category: str = "labels") -> None:
    assert labels is not None, f"Model is in training model but {category} are not passed"


class ModelOutput(OrderedDict):
    """torchmultimodal/utils/common.py:122-139: a dataclass that also behaves like a read-only mapping of its fields."""

    def keys(self) -> Any:
        for f in fields(self):  # type: ignore
            yield f.name

    def __getitem__(self, key: Any) -> Any:
        return getattr(self, key)

    def __iter__(self) -> Any:
        yield from self.keys()

    def values(self) -> Any:
        for f in fields(self):  # type: ignore
            yield getattr(self, f.name)

    def items(self) -> Any:
        for f in fields(self):  # type: ignore
            yield f.name, getattr(self, f.name)


def _output(name: str, required=(), optional=(), base=ModelOutput, extra=None):
    """Builds one of the reference's output containers (losses/flava.py:30-81): a dataclass over `base` whose fields are
    `required` (positional tensors) followed by `optional` (default None) and `extra` (name -> default factory)."""
    spec = [(f, Tensor) for f in required]
    spec += [(f, Optional[Any], field(default_factory=fac)) for f, fac in (extra or {}).items()]
    spec += [(f, Optional[Any], None) for f in optional]
    cls = make_dataclass(name, spec, bases=(base,))
    cls.__module__ = __name__
    return cls


_HEAD_FIELDS = ("logits", "loss")
ITMLossOutput = _output("ITMLossOutput", _HEAD_FIELDS)
MaskedPredictionLossOutput = _output("MaskedPredictionLossOutput", _HEAD_FIELDS)
FLAVAGlobalContrastiveLossOutput = _output(
    "FLAVAGlobalContrastiveLossOutput",
    ("text_embedding", "image_embedding", "logit_scale", "image_logits", "text_logits", "image_loss", "text_loss", "loss"),
    base=OrderedDict)
_LOSS_NAMES = ("mmm_text_loss", "mmm_image_loss", "mim_loss", "mlm_loss", "itm_loss", "global_contrastive_loss")
FLAVAPretrainingLossesCollection = _output("FLAVAPretrainingLossesCollection", optional=_LOSS_NAMES)
FLAVAPretrainingLossOutput = _output(
    "FLAVAPretrainingLossOutput",
    optional=("mlm_output", "mim_output", "mmm_text_output", "mmm_image_output", "itm_output", "global_contrastive_output",
              "image_sequence", "text_sequence", "image_masked_sequence", "text_masked_sequence", "multimodal_sequence",
              "multimodal_masked_sequence"),
    extra={"losses": FLAVAPretrainingLossesCollection})


# ---------------------------------------------------------------------------------------------------------------------
# runtime helpers (kernel launches only)
# ---------------------------------------------------------------------------------------------------------------------
def _wants_graph(*mods_or_tensors: Any) -> bool:
    """True when the caller expects an autograd graph through this call: grad mode on and a trainable parameter or an
    input that requires grad.  Those calls take the autograd Functions of engine_flava_heads / engine_flava_train (same
    kernels forward, explicit backward schedules); everything else runs the forward-value path under no_grad."""
    if not torch.is_grad_enabled():
        return False
    for x in mods_or_tensors:
        if isinstance(x, nn.Module):
            if any(p.requires_grad for p in x.parameters()):
                return True
        elif isinstance(x, Tensor) and x.requires_grad:
            return True
    return False


def _no_graph(what: str, *mods_or_tensors: Any) -> None:
    """Forward-value-only entry points: refuse to hand back a value that autograd would treat as a constant."""
    if _wants_graph(*mods_or_tensors):
        raise MMBError(f"{what} computes forward values only on this call path (no labels -> no loss to differentiate); "
                       "call it under torch.no_grad()")


def _rows_view(x: Tensor) -> Tensor:
    """[B, S, d] fp32 (possibly a slice along S of a larger buffer) -> something whose rows are d contiguous floats."""
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(-1) != 1:
        x = x.contiguous()
    return x


def _flat_row_index(x: Tensor, keep: Tensor) -> (Tensor, Tensor, int):
    """Flat row numbers (in units of x's row pitch, relative to x's storage start of its first element) of x[b, s, :]
    for every kept (b, s).  x: [B, S, d] with x.stride(-1) == 1 and x.stride(0) a multiple of x.stride(1)."""
    B, S, d = x.shape
    ld = x.stride(1)
    if x.stride(0) % ld != 0:
        x = x.contiguous()
        ld = d
    per_b = x.stride(0) // ld
    bs = keep.nonzero(as_tuple=False)                      # [n, 2] (b, s): index bookkeeping on the labels, not the path
    idx = (bs[:, 0] * per_b + bs[:, 1]).contiguous()
    return x, idx, ld


def _bf16(t: Tensor) -> Tensor:
    return ops.cast_bf16(t.detach().contiguous())


def _select_rows_bf16(x: Tensor, keep: Tensor) -> Tensor:
    x = _rows_view(x)
    x, idx, _ = _flat_row_index(x, keep)
    out = torch.empty((max(int(idx.numel()), 1), x.shape[-1]), device=x.device, dtype=torch.bfloat16)
    if idx.numel() == 0:
        out.zero_()
        return out[:0]
    ops.gather_rows_idx_cast(x, idx, out, x.shape[-1])
    return out


def _linear_f32(a_bf16: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """fp32 [M, N] = a @ weight^T + bias on the tensor cores; N padded to a multiple of 8 internally (TMA row pitch)."""
    M, N = a_bf16.shape[0], weight.shape[0]
    Np = (N + 7) // 8 * 8
    w = _bf16(weight)
    b = bias.detach().float().contiguous() if bias is not None else None
    if Np != N:
        wp = torch.zeros((Np, weight.shape[1]), device=w.device, dtype=torch.bfloat16)
        wp[:N].copy_(w)
        w = wp
        if b is not None:
            bp = torch.zeros(Np, device=w.device, dtype=torch.float32)
            bp[:N].copy_(b)
            b = bp
    out = torch.empty((M, Np), device=a_bf16.device, dtype=torch.float32)
    ops.gemm(a_bf16, w, epilogue=ops.EPI_F32, bias=b, out=out)
    return out[:, :N]


def _cross_entropy(logits: Tensor, labels: Tensor, ignore_index: int) -> Tensor:
    """nn.CrossEntropyLoss(ignore_index)(logits, labels): mean over the kept rows (NaN when none is kept, as torch)."""
    M, V = logits.shape
    accum = torch.zeros(2, device=logits.device, dtype=torch.float32)
    ops.ce_labels(logits, labels.contiguous(), 1, ignore_index, M, V, None, accum)
    return accum[0] / accum[1]


class Pooler(nn.Module):
    def __init__(self, hidden_size: int = 768, **kwargs: Any):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states: Tensor) -> Tensor:
        # losses/flava.py:91-97: dense + tanh on the first token
        if _wants_graph(self, hidden_states):
            from ...engine_flava_train import first_token_linear
            return first_token_linear(hidden_states, self.dense, use_tanh=True)
        x = _rows_view(hidden_states)
        keep = torch.zeros(x.shape[:2], dtype=torch.bool, device=x.device)
        keep[:, 0] = True
        first = _select_rows_bf16(x, keep)
        pooled = _linear_f32(first, self.dense.weight, self.dense.bias).contiguous()
        return ops.tanh_(pooled)


class TwoWayHead(nn.Module):
    def __init__(self, hidden_size: int = 768, **kwargs: Any):
        super().__init__()
        self.seq_relationship = nn.Linear(hidden_size, 2)

    def forward(self, pooled_output: Tensor) -> Tensor:
        if _wants_graph(self, pooled_output):
            from ...engine_flava_heads import small_linear
            return small_linear(pooled_output, self.seq_relationship)
        return _linear_f32(_bf16(pooled_output.float()), self.seq_relationship.weight, self.seq_relationship.bias)


class ITMLoss(nn.Module):
    def __init__(self, hidden_size: int = 768, ignore_index: int = -1, **kwargs: Any):
        super().__init__()
        self.pooler = Pooler(hidden_size=hidden_size)
        self.cls = TwoWayHead(hidden_size=hidden_size)
        self.ce_loss = nn.CrossEntropyLoss(ignore_index=ignore_index)
        self.ignore_index = ignore_index

    def forward(self, hidden_states: Tensor, labels: Tensor) -> ITMLossOutput:
        if self.training:
            assert_labels_are_present(labels, "itm labels")
        pooled_output = self.pooler(hidden_states)
        scores = self.cls(pooled_output)
        if labels is None:
            loss = torch.zeros((), device=scores.device)
        elif scores.requires_grad:
            from ...engine_flava_heads import cross_entropy
            loss = cross_entropy(scores.reshape(-1, 2), labels.reshape(-1).long(), self.ignore_index)
        else:
            loss = _cross_entropy(scores.reshape(-1, 2), labels.reshape(-1).long(), self.ignore_index)
        return ITMLossOutput(logits=scores, loss=loss)


class MaskedPredictionHead(nn.Module):
    def __init__(self, hidden_size: int = 768, vocab_size: int = 30522,
                 transform_act_fn: Callable[[Tensor], Tensor] = nn.functional.gelu, layer_norm_eps: float = 1e-5,
                 use_fp32_layer_norm: bool = True, **kwargs: Any):
        super().__init__()
        if transform_act_fn is not nn.functional.gelu:
            raise NotImplementedError("MaskedPredictionHead: only the erf-GELU transform (the reference default) has a "
                                      "GEMM epilogue")
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.transform_act_fn = transform_act_fn
        self.layer_norm: nn.LayerNorm
        if use_fp32_layer_norm:
            self.layer_norm = Fp32LayerNorm(hidden_size, eps=layer_norm_eps)
        else:
            self.layer_norm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        # output weights + an output-only bias, linked as in the reference (:164-172)
        self.decoder = nn.Linear(hidden_size, vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(vocab_size))
        self.decoder.bias = self.bias

    def _forward_rows(self, rows_bf16: Tensor) -> Tensor:
        """rows_bf16 [n, hidden] -> fp32 logits [n, vocab]."""
        n, hdim = rows_bf16.shape
        if n == 0:
            return torch.empty((0, self.decoder.weight.shape[0]), device=rows_bf16.device, dtype=torch.float32)
        _, act = ops.gemm(rows_bf16, _bf16(self.dense.weight), bias=self.dense.bias.detach().float().contiguous(),
                          epilogue=ops.EPI_BF16_ACT, act=ops.ACT_GELU_ERF)
        ln = torch.empty((n, hdim), device=rows_bf16.device, dtype=torch.bfloat16)
        ops.add_layernorm_fwd(None, act, None, ln, None, self.layer_norm.weight.detach().float().contiguous(),
                              self.layer_norm.bias.detach().float().contiguous(), None, None, n, hdim,
                              self.layer_norm.eps)
        return _linear_f32(ln, self.decoder.weight, self.bias)

    def forward(self, hidden_states: Tensor) -> Tensor:
        _no_graph("MaskedPredictionHead", self, hidden_states)
        x = _rows_view(hidden_states)
        lead = x.shape[:-1]
        x3 = x.reshape(-1, 1, x.shape[-1]) if x.dim() != 3 else x
        keep = torch.ones(x3.shape[:2], dtype=torch.bool, device=x.device)
        return self._forward_rows(_select_rows_bf16(x3, keep)).reshape(*lead, -1)


class MaskedPredictionLoss(nn.Module):
    def __init__(self, hidden_size: int = 768, vocab_size: int = 30522,
                 transform_act_fn: Callable[[Tensor], Tensor] = nn.functional.gelu, layer_norm_eps: float = 1e-5,
                 ignore_index: int = -1, ignore_nan: bool = False, **kwargs: Any):
        super().__init__()
        self.cls = MaskedPredictionHead(hidden_size=hidden_size, vocab_size=vocab_size,
                                        transform_act_fn=transform_act_fn, layer_norm_eps=layer_norm_eps)
        self.ignore_index = ignore_index
        self.vocab_size = vocab_size
        self.ce_loss = nn.CrossEntropyLoss(ignore_index=ignore_index)
        self.ignore_nan = ignore_nan

    def forward(self, hidden_states: Tensor, masked_labels: Optional[Tensor] = None) -> MaskedPredictionLossOutput:
        if self.training:
            assert_labels_are_present(masked_labels, "masked labels")
        training = _wants_graph(self, hidden_states)
        if masked_labels is not None and training:
            from ...engine_flava_heads import masked_prediction
            masked_tokens = masked_labels.ne(self.ignore_index)                  # :212-215
            kept_labels = masked_labels[masked_tokens].long()
            prediction, masked_loss = masked_prediction(hidden_states, masked_tokens, kept_labels, self.cls,
                                                        self.ignore_index)
        elif masked_labels is not None:
            masked_tokens = masked_labels.ne(self.ignore_index)                  # :212-215
            kept_labels = masked_labels[masked_tokens].long()
            prediction = self.cls._forward_rows(_select_rows_bf16(hidden_states, masked_tokens))
            if kept_labels.numel() == 0:
                masked_loss = torch.full((), float("nan"), device=hidden_states.device)   # CE mean over zero rows
            else:
                masked_loss = _cross_entropy(prediction, kept_labels, self.ignore_index)
        else:
            prediction = self.cls(hidden_states)
            masked_loss = torch.zeros((), device=hidden_states.device)
        if self.ignore_nan and torch.isnan(masked_loss):   # the host read-back only when the option is on
            warnings.warn("NaN detected in masked_loss. Replacing it with 0.")
            masked_loss = torch.nan_to_num(masked_loss, nan=0.0)
        return MaskedPredictionLossOutput(logits=prediction, loss=masked_loss)


class FLAVAGlobalContrastiveLoss(nn.Module):
    def __init__(self, logit_scale: Union[float, nn.Parameter] = None, image_embedding_size: int = 768,
                 text_embedding_size: int = 768, projection_size: int = 768, image_embedding_index: int = 0,
                 text_embedding_index: int = 0):
        super().__init__()
        if logit_scale is None:
            logit_scale = math.log(1 / 0.07)
        if isinstance(logit_scale, nn.Parameter):
            self.logit_scale = logit_scale
        else:
            self.logit_scale = nn.Parameter(logit_scale * torch.ones([]))

    @staticmethod
    def _normalize(x: Tensor) -> Tensor:
        x = x.detach().float().contiguous()
        B, E = x.shape
        y = torch.empty_like(x)
        ops.l2norm_fwd(x, y, None, torch.empty(B, device=x.device), B, E)
        return y

    def forward(self, image_sequence: Tensor, text_sequence: Tensor, mask: Tensor) -> FLAVAGlobalContrastiveLossOutput:
        training = _wants_graph(self, image_sequence, text_sequence)
        if training:
            from ...autograd import l2_normalize
            text_embedding = l2_normalize(text_sequence.float())
            image_embedding = l2_normalize(image_sequence.float())
        else:
            text_embedding = self._normalize(text_sequence)
            image_embedding = self._normalize(image_sequence)
        self.logit_scale.data.clamp_(0, 4.6052)                                  # :273
        with torch.set_grad_enabled(training):
            output = contrastive_loss_with_temperature(
                embeddings_a=image_embedding, embeddings_b=text_embedding, logit_scale=self.logit_scale, mask=mask,
                backprop_type=BackpropType.GLOBAL)                               # always GLOBAL for FLAVA (:280-281)
        return FLAVAGlobalContrastiveLossOutput(
            loss=output.loss, image_logits=output.logits_a, text_logits=output.logits_b, image_loss=output.loss_a,
            text_loss=output.loss_b, text_embedding=text_embedding, image_embedding=image_embedding,
            logit_scale=self.logit_scale.data)


class FLAVAPretrainingLoss(nn.Module):
    def __init__(self, logit_scale: Union[float, nn.Parameter] = None, hidden_size: int = 768,
                 text_vocab_size: int = 30522, image_vocab_size: int = 8192,
                 transform_act_fn: Callable[[Tensor], Tensor] = nn.functional.gelu, layer_norm_eps: float = 1e-5,
                 ignore_index: int = -1, mlm_weight: float = 1.0, mim_weight: float = 1.0,
                 contrastive_loss_weight: float = 1.0, mmm_image_loss_weight: float = 1.0,
                 mmm_text_loss_weight: float = 1.0, itm_loss_weight: float = 1.0, **kwargs: Any):
        super().__init__()
        # module creation order = the reference's (:316-356): identical parameters under the same seed
        self.contrastive_loss = FLAVAGlobalContrastiveLoss(
            logit_scale=logit_scale, image_embedding_size=hidden_size, text_embedding_size=hidden_size,
            projection_size=hidden_size)
        mk = dict(hidden_size=hidden_size, transform_act_fn=transform_act_fn, layer_norm_eps=layer_norm_eps,
                  ignore_index=ignore_index)
        self.mlm_loss = MaskedPredictionLoss(vocab_size=text_vocab_size, **mk)
        self.mim_loss = MaskedPredictionLoss(vocab_size=image_vocab_size, **mk)
        self.mmm_loss = nn.ModuleDict({"mlm": MaskedPredictionLoss(vocab_size=text_vocab_size, **mk),
                                       "mim": MaskedPredictionLoss(vocab_size=image_vocab_size, **mk)})
        self.itm_loss = ITMLoss(hidden_size=hidden_size, ignore_index=ignore_index)
        self.mim_weight = mim_weight
        self.mlm_weight = mlm_weight
        self.contrastive_loss_weight = contrastive_loss_weight
        self.mmm_image_loss_weight = mmm_image_loss_weight
        self.mmm_text_loss_weight = mmm_text_loss_weight
        self.itm_loss_weight = itm_loss_weight

    def _weighted(self, out, weight: float, slot: str, result: FLAVAPretrainingLossOutput):
        """`outputs.<x>_output.loss *= weight; outputs.losses.<x>_loss = ...` of the reference, without the in-place op."""
        out.loss = out.loss * weight
        setattr(result, slot + "_output", out)
        setattr(result.losses, slot + "_loss", out.loss)

    def _positive_pairs(self, mm_masked: Tensor, itm_labels: Optional[Tensor]) -> Tensor:
        # :416-424 — pairs with itm label != 0; if there is none, every pair counts
        n = mm_masked.size(0)
        if itm_labels is None:
            return torch.ones(n, dtype=torch.bool, device=mm_masked.device)
        pos = itm_labels.ne(0)
        return pos if bool(pos.any()) else torch.ones_like(pos)

    def forward(self, image_sequence: Optional[Tensor] = None, text_sequence: Optional[Tensor] = None,
                image_masked_sequence: Optional[Tensor] = None, text_masked_sequence: Optional[Tensor] = None,
                multimodal_sequence: Optional[Tensor] = None, multimodal_masked_sequence: Optional[Tensor] = None,
                itm_labels: Optional[Tensor] = None, mim_labels: Optional[Tensor] = None,
                mlm_labels: Optional[Tensor] = None, projected_image_embeddings: Optional[Tensor] = None,
                projected_text_embeddings: Optional[Tensor] = None) -> FLAVAPretrainingLossOutput:
        """Same branches as the reference (:370-484).  Every tensor op on hidden states is a kernel of this library; the
        boolean / index bookkeeping on the (tiny) label tensors stays in torch."""
        res = FLAVAPretrainingLossOutput()
        mm = multimodal_masked_sequence
        unimodal = mm is None
        pos_mask = None

        # -- unimodal masked prediction (:386-413): the sequence tail that lines up with the labels (CLS dropped) --
        if unimodal and image_masked_sequence is not None and self.mim_weight > 0:
            tail = mim_labels.size(1) if mim_labels is not None else image_masked_sequence.size(1) - 1
            self._weighted(self.mim_loss(image_masked_sequence[:, -tail:, :], mim_labels), self.mim_weight, "mim", res)
        if unimodal and text_masked_sequence is not None and self.mlm_weight > 0:
            tail = mlm_labels.size(1) if mlm_labels is not None else text_masked_sequence.size(1) - 1
            self._weighted(self.mlm_loss(text_masked_sequence[:, -tail:, :], mlm_labels), self.mlm_weight, "mlm", res)

        # -- image-text matching + restriction of the masked-multimodal losses to the positive pairs (:415-435) --
        if not unimodal and self.itm_loss_weight > 0:
            pos_mask = self._positive_pairs(mm, itm_labels)
            self._weighted(self.itm_loss(mm, itm_labels), self.itm_loss_weight, "itm", res)
            negatives = ~pos_mask
            if bool(negatives.any()):
                if mlm_labels is None or mim_labels is None:
                    # label-free (inference) call: predictions cover the kept sequences only, as `mm[pos_mask]` does
                    mm = mm[pos_mask]
                    mlm_labels = None if mlm_labels is None else mlm_labels[pos_mask]
                    mim_labels = None if mim_labels is None else mim_labels[pos_mask]
                else:
                    # instead of copying the kept sequences, the labels of the dropped pairs become ignore_index: the
                    # masked-prediction heads then select exactly the rows `mm[pos_mask]` would have contributed
                    ign_t, ign_i = self.mmm_loss["mlm"].ignore_index, self.mmm_loss["mim"].ignore_index
                    mlm_labels = mlm_labels.masked_fill(negatives[:, None], ign_t)
                    mim_labels = mim_labels.masked_fill(negatives[:, None], ign_i)

        # -- masked multimodal modelling (:437-466): text tokens are the sequence tail, image patches follow the two CLS --
        if not unimodal and self.mmm_text_loss_weight > 0:
            tail = mlm_labels.size(1) if mlm_labels is not None else text_masked_sequence.size(1) - 1
            self._weighted(self.mmm_loss["mlm"](mm[:, -tail:, :], mlm_labels), self.mmm_text_loss_weight, "mmm_text", res)
        if not unimodal and self.mmm_image_loss_weight > 0:
            # reference quirk kept (:454-458): the patch count comes from mim_labels only when mlm_labels is given
            n_patch = mim_labels.size(1) if mlm_labels is not None else image_masked_sequence.size(1) - 1
            self._weighted(self.mmm_loss["mim"](mm[:, 2:2 + n_patch, :], mim_labels), self.mmm_image_loss_weight,
                           "mmm_image", res)

        # -- global contrastive loss over the projected CLS embeddings (:468-482) --
        have_proj = projected_image_embeddings is not None and projected_text_embeddings is not None
        if have_proj and self.contrastive_loss_weight > 0:
            gc = self.contrastive_loss(projected_image_embeddings, projected_text_embeddings, pos_mask)
            gc.loss = gc.loss * self.contrastive_loss_weight
            res.global_contrastive_output = gc
            res.losses.global_contrastive_loss = gc.loss
        return res
