"""CLIP text transform (text half): the oracle (oracle/clip_text_oracle.py) and the native BPE encoder
behind the drop-in modules (host code in libmmb200.so — no GPU involved) against ids produced by the unmodified reference
classes (tests/golden/clip_text_golden.pt), bit-exact: integer work."""
import os

import pytest
import torch

from oracle import clip_text_oracle as TO

HERE = os.path.dirname(__file__)
MERGES = os.path.join(HERE, "golden", "clip_bpe_merges.bpe")
GOLD = os.path.join(HERE, "golden", "clip_text_golden.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


@pytest.mark.parametrize("nm", [None, 40])
def test_oracle_and_native_encoder_reproduce_reference_ids(gold, nm):
    from multimodal_b200.transforms.clip_text_transform import CLIPBPETransform

    merges_text = open(MERGES, encoding="utf-8").read()
    tok = TO.Tokenizer(merges_text, num_merges=nm)
    native = CLIPBPETransform(MERGES, num_merges=nm)
    assert native.vocab_size == gold[f"vocab_size.{nm}"] == len(tok.ids)
    for text, want in zip(gold["texts"], gold[f"encode.{nm}"]):
        assert tok.encode(text) == want, text
        assert native(text) == want, text
    assert native(gold["texts"]) == gold[f"encode.{nm}"]            # batch call, word cache warm
    assert native.token_id("<|endoftext|>") == native.vocab_size - 1
    with pytest.raises(KeyError):
        native.token_id("no such symbol")


@pytest.mark.parametrize("L,pad", [(12, None), (77, None), (16, "the")])
def test_text_transform_reproduces_reference_tensors(gold, L, pad):
    from multimodal_b200.transforms.clip_text_transform import CLIPTextTransform

    t = CLIPTextTransform(text_max_length=L, text_bpe_merges_path=MERGES, num_merges=None, text_pad_token=pad, device="cpu")
    nonempty = [x for x in gold["texts"] if x.strip()]
    got = t(nonempty)
    want = gold[f"transform.{L}.{pad}"]
    assert got.dtype == torch.int64 and torch.equal(got, want)
    assert torch.equal(t(nonempty[1]), gold[f"transform_single.{L}.{pad}"])
    assert torch.equal(t(["hello", "a dog"]), gold[f"transform_short_batch.{L}.{pad}"])
    tok = TO.Tokenizer(open(MERGES, encoding="utf-8").read())
    assert TO.text_transform(tok, nonempty, L, pad) == want.tolist()


def test_random_texts_native_equals_oracle():
    """Size-independent property: on random strings over a mixed alphabet the native encoder and the oracle agree, and
    every id is a valid vocabulary index; repeated calls (cache hits) return the same ids."""
    import random

    from multimodal_b200.transforms.clip_text_transform import CLIPBPETransform

    rng = random.Random(0)
    alphabet = "abcdefghijklmnopqrstuvwxyzABC 0123456789'.,!?-éüß日本<|>/\t"
    tok = TO.Tokenizer(open(MERGES, encoding="utf-8").read())
    native = CLIPBPETransform(MERGES)
    texts = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 60))) for _ in range(300)]
    want = [tok.encode(t) for t in texts]
    assert native(texts) == want
    assert native(texts) == want
    assert all(0 <= i < native.vocab_size for ids in want for i in ids)


def test_default_merges_path_needs_a_local_file():
    from multimodal_b200.transforms.clip_text_transform import CLIPTextTransform

    with pytest.raises(NotImplementedError):
        CLIPTextTransform(device="cpu")
