"""Generates tests/golden/clip_text_golden.pt: token ids produced by the UNMODIFIED reference classes
(`CLIPBPETokenizer`, `CLIPTextTransform`, torchmultimodal/transforms/clip_transform.py:82-298) on the synthetic merges
file tests/golden/clip_bpe_merges.bpe (the real `clip_merges.bpe` is a download).  Two stand-ins make the reference
module importable offline: `ftfy` (imported at module level, used only by `basic_clean`, which `encode` never calls) is
an identity stub, and `iopath` is oracle/iopath_shim (local files pass through).
Run: python tests/golden/make_clip_text_golden.py"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "iopath_shim"))
sys.path.insert(0, "/root/reference")
_ftfy = types.ModuleType("ftfy")
_ftfy.fix_text = lambda t: t
sys.modules.setdefault("ftfy", _ftfy)

TEXTS = [
    "A photo of a dog, running in the park!",
    "hello world",
    "  The CATS are playing; it's 2023 — don't stop...  ",
    "café naïve <|endoftext|> x",
    "",
    "a photo of cats playing in the park with dogs and more cats and dogs running in the park hello world the end",
    "ÀÉÎ õü ß 12ab3 'll 'RE ???!!! 日本語 text",
    "<|startoftext|>hello<|endoftext|>",
]


def main():
    from torchmultimodal.transforms.clip_transform import CLIPBPETokenizer, CLIPTextTransform

    merges = os.path.join(HERE, "clip_bpe_merges.bpe")
    out = {"texts": TEXTS}
    for nm in (None, 40):
        tok = CLIPBPETokenizer(merges, num_merges=nm)
        out[f"encode.{nm}"] = [tok.encode(t) for t in TEXTS]
        out[f"vocab_size.{nm}"] = tok.vocab_size
    nonempty = [t for t in TEXTS if t.strip()]
    for L, pad in ((12, None), (77, None), (16, "the")):
        tt = CLIPTextTransform(text_max_length=L, text_bpe_merges_path=merges, num_merges=None, text_pad_token=pad)
        out[f"transform.{L}.{pad}"] = tt(nonempty)
        out[f"transform_single.{L}.{pad}"] = tt(nonempty[1])
        out[f"transform_short_batch.{L}.{pad}"] = tt(["hello", "a dog"])
    path = os.path.join(HERE, "clip_text_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
