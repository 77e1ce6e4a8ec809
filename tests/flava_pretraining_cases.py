"""Shared definitions of the FLAVA pre-training-loss parity cases: seeded loss module, seeded hidden-state
sequences and labels, and a deterministic stand-in for the image codebook.  Used by the fixture generator
(tests/golden/make_flava_pretraining_golden.py, which feeds them to the REFERENCE classes) and by the parity tests."""
import torch
from torch import nn

import flava_cases as FC

LOSS_KW = dict(hidden_size=128, text_vocab_size=1002, image_vocab_size=512)   # 1002: not a multiple of 8 (padded GEMM N)
B, S_IMG, S_TXT = 5, 17, 12          # 16 patches + CLS, 12 text tokens, multimodal = 1 + 17 + 12 = 30


def build_loss(cls):
    """cls = reference or multimodal_b200 FLAVAPretrainingLoss: same seed, same creation order -> identical weights."""
    torch.manual_seed(0)
    m = cls(**LOSS_KW)
    g = torch.Generator().manual_seed(13)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() > 0:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    return m.eval()


def loss_inputs():
    g = torch.Generator().manual_seed(7)
    d = LOSS_KW["hidden_size"]
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    mlm = torch.full((B, S_TXT), -1, dtype=torch.long)
    mim = torch.full((B, S_IMG - 1), -1, dtype=torch.long)
    pick_t = torch.rand(B, S_TXT, generator=g) < 0.3
    pick_i = torch.rand(B, S_IMG - 1, generator=g) < 0.4
    pick_t[0, 1] = True
    pick_i[1, 3] = True
    mlm[pick_t] = torch.randint(0, LOSS_KW["text_vocab_size"], (int(pick_t.sum()),), generator=g)
    mim[pick_i] = torch.randint(0, LOSS_KW["image_vocab_size"], (int(pick_i.sum()),), generator=g)
    return dict(
        image_sequence=r(B, S_IMG, d), text_sequence=r(B, S_TXT, d), image_masked_sequence=r(B, S_IMG, d),
        text_masked_sequence=r(B, S_TXT, d), multimodal_masked_sequence=r(B, 1 + S_IMG + S_TXT, d),
        itm_labels=torch.tensor([1, 0, 1, 1, 0]), mlm_labels=mlm, mim_labels=mim,
        projected_image_embeddings=r(B, 64), projected_text_embeddings=r(B, 64))


def loss_calls():
    """name -> kwargs of FLAVAPretrainingLoss.forward: the unimodal branch (MIM + MLM + unmasked contrastive) and the
    multimodal branch (ITM + MMM text / image + contrastive over the positive pairs)."""
    x = loss_inputs()
    uni = {k: x[k] for k in ("image_sequence", "text_sequence", "image_masked_sequence", "text_masked_sequence",
                             "mlm_labels", "mim_labels", "projected_image_embeddings", "projected_text_embeddings")}
    return {"unimodal": uni, "multimodal": dict(x)}


def flatten_loss_output(o):
    out = {}
    for k, v in o.losses.items():
        if v is not None:
            out[f"losses.{k}"] = v.detach().float().cpu()
    for name in ("mlm_output", "mim_output", "mmm_text_output", "mmm_image_output", "itm_output"):
        v = getattr(o, name)
        if v is not None:
            out[f"{name}.logits"] = v.logits.detach().float().cpu()
            out[f"{name}.loss"] = v.loss.detach().float().cpu()
    gc = o.global_contrastive_output
    if gc is not None:
        for k in ("loss", "image_loss", "text_loss", "image_logits", "text_logits", "image_embedding", "text_embedding"):
            out[f"global_contrastive_output.{k}"] = getattr(gc, k).detach().float().cpu()
    return out


# ---- whole model: FLAVAForPreTraining(flava_model, stub codebook, loss) -----------------------------------------------
MODEL_KW = dict(image_hidden_size=128, image_num_attention_heads=2, image_num_hidden_layers=1, image_intermediate_size=256,
                image_size=32, patch_size=8, text_hidden_size=128, text_num_attention_heads=2, text_num_hidden_layers=1,
                text_intermediate_size=256, vocab_size=100, max_position_embeddings=32, multimodal_hidden_size=128,
                multimodal_num_attention_heads=2, multimodal_num_hidden_layers=1, multimodal_intermediate_size=256,
                text_and_image_proj_size=64)
MODEL_LOSS_KW = dict(hidden_size=128, text_vocab_size=100, image_vocab_size=64)
MODEL_B, MODEL_ST = 4, 12


class StubCodebook(nn.Module):
    """Deterministic stand-in for DalleVAEEncoder (DALL_E package + download): token id per patch from the patch mean."""

    def __init__(self, patch: int, vocab: int):
        super().__init__()
        self.patch, self.vocab = patch, vocab

    def forward(self, img):
        m = torch.nn.functional.avg_pool2d(img.float().mean(1, keepdim=True), self.patch).flatten(1)
        return (m * 1000.0).abs().long() % self.vocab


def build_model(flava_model, FLAVAForPreTraining, FLAVAPretrainingLoss):
    torch.manual_seed(0)
    model = flava_model(**MODEL_KW)
    loss = FLAVAPretrainingLoss(**MODEL_LOSS_KW)
    m = FLAVAForPreTraining(model=model, image_codebook=StubCodebook(MODEL_KW["patch_size"], MODEL_LOSS_KW["image_vocab_size"]),
                            loss=loss)
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() > 0:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    return m.eval()


def model_inputs():
    g = torch.Generator().manual_seed(23)
    kw = MODEL_KW
    image = torch.randn(MODEL_B, 3, kw["image_size"], kw["image_size"], generator=g)
    text = torch.randint(1, kw["vocab_size"] - 1, (MODEL_B, MODEL_ST), generator=g)
    text[1, MODEL_ST - 3:] = 0
    text_masked = text.clone()
    mlm_labels = torch.full((MODEL_B, MODEL_ST), -1, dtype=torch.long)
    for b in range(MODEL_B):
        for s in (2, 5):
            mlm_labels[b, s] = text[b, s]
            text_masked[b, s] = kw["vocab_size"] - 1
    P = (kw["image_size"] // kw["patch_size"]) ** 2
    side = kw["image_size"] // kw["patch_size"]
    patches_mask = (torch.rand(MODEL_B, side, side, generator=g) < 0.4)
    patches_mask[:, 0, 0] = True
    return dict(image=image, text=text, image_for_codebook=image.clone(), image_patches_mask=patches_mask,
                text_masked=text_masked, itm_labels=torch.tensor([1, 1, 0, 1]), mlm_labels=mlm_labels,
                skip_unmasked_mm_encoder=False), P


param_checksum = FC.param_checksum
