"""`CLIPBPETransform`, `CLIPTextTransform`, `CLIPTransform` — drop-ins for the text half of
torchmultimodal/transforms/clip_transform.py:82-298, 355-420.

Tokenisation is host string processing (there is nothing to move to the GPU); what changes is where the time goes: the
byte-pair merge loop runs in native code for a whole batch at once (`mmb_bpe_encode`, word cache inside the encoder),
Python only lower-cases, regex-splits (`regex`, the same pattern as the reference) and assembles the [B, L] id tensor —
start / end tokens, truncation to L - 2, the reference's two-stage padding — which is uploaded once.
`text_bpe_merges_path` must be a local file (the reference's default is a download URL; no network here).  `ftfy` text
repair is applied by the reference's `basic_clean` helper only, which its `encode` path never calls, so ids do not depend on it.
"""
import ctypes
import weakref
from typing import List, Optional, Tuple, Union

import torch
from torch import nn, Tensor

from .. import _lib
from .clip_transform import CLIP_DEFAULT_MEAN, CLIP_DEFAULT_STD, CLIPImageTransform

CLIP_DEFAULT_VOCAB_BPE_PATH = "http://download.pytorch.org/models/text/clip_merges.bpe"
_PATTERN = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""


def _destroy(lib, handle) -> None:
    lib.mmb_bpe_destroy(handle)


def _read_merges(path: Optional[str]) -> bytes:
    if path is None or str(path).startswith(("http://", "https://")):
        raise NotImplementedError("pass text_bpe_merges_path=<local clip_merges.bpe>: the reference downloads "
                                  f"{CLIP_DEFAULT_VOCAB_BPE_PATH}, and there is no network access here")
    with open(path, "r", encoding="utf-8") as f:      # text mode: universal newlines, as the reference reads it
        return f.read().encode("utf-8")


class CLIPBPETransform(nn.Module):
    """Byte-level BPE encoder (clip_transform.py:200-241): `forward(str | List[str]) -> List[int] | List[List[int]]`."""

    def __init__(self, bpe_path: Optional[str] = CLIP_DEFAULT_VOCAB_BPE_PATH, bos_token: Optional[str] = "<|startoftext|>",
                 eos_token: Optional[str] = "<|endoftext|>", num_merges: Optional[int] = None):
        super().__init__()
        import regex

        data = _read_merges(bpe_path)
        self._pat = regex.compile(_PATTERN, regex.IGNORECASE)
        h, n = ctypes.c_void_p(), ctypes.c_int()
        _lib.check(_lib.lib().mmb_bpe_create(data, len(data), int(num_merges or 0), bos_token.encode("utf-8"),
                                             eos_token.encode("utf-8"), ctypes.byref(h), ctypes.byref(n)), "mmb_bpe_create")
        self._h, self.vocab_size = h, int(n.value)
        self.bos_token, self.eos_token = bos_token, eos_token
        weakref.finalize(self, _destroy, _lib.lib(), h)     # frees the native encoder with the module

    def token_id(self, token: str) -> int:
        i = int(_lib.lib().mmb_bpe_token_id(self._h, token.encode("utf-8")))
        if i < 0:
            raise KeyError(token)
        return i

    def encode_batch(self, texts: List[str]) -> List[List[int]]:
        pieces: List[bytes] = []
        per_text: List[int] = []
        for t in texts:
            found = self._pat.findall(t.lower().strip())          # clip_transform.py:177-179
            per_text.append(len(found))
            pieces.extend(p.encode("utf-8") for p in found)
        n = len(pieces)
        if n == 0:
            return [[] for _ in texts]
        offs = (ctypes.c_longlong * (n + 1))()
        total = 0
        for i, p in enumerate(pieces):
            offs[i] = total
            total += len(p)
        offs[n] = total
        ids, counts = (ctypes.c_int * max(total, 1))(), (ctypes.c_int * n)()
        _lib.check(_lib.lib().mmb_bpe_encode(self._h, b"".join(pieces), offs, n, ids, counts), "mmb_bpe_encode")
        out, w, k = [], 0, 0
        for cnt in per_text:
            m = sum(counts[k:k + cnt])
            out.append(list(ids[w:w + m]))
            w += m
            k += cnt
        return out

    def forward(self, text: Union[str, List[str]]) -> Union[List[int], List[List[int]]]:
        if isinstance(text, str):
            return self.encode_batch([text])[0]
        return self.encode_batch(list(text))


class CLIPTextTransform(nn.Module):
    """clip_transform.py:244-298: BPE ids, truncated to `text_max_length - 2`, wrapped in start / end tokens, padded to
    `text_max_length`; int64 [L] for a string, [B, L] for a list.  `device`: where the id tensor is returned (the
    reference returns a CPU tensor; the towers want it on the GPU)."""

    def __init__(self, text_max_length: int = 77, text_start_token: str = "<|startoftext|>",
                 text_end_token: str = "<|endoftext|>", text_pad_token: Optional[str] = None,
                 text_bpe_merges_path: str = CLIP_DEFAULT_VOCAB_BPE_PATH, num_merges: Optional[int] = 48894,
                 device: Union[str, torch.device] = "cuda") -> None:
        super().__init__()
        self.tokenizer = CLIPBPETransform(text_bpe_merges_path, text_start_token, text_end_token, num_merges)
        self.start_id = self.tokenizer([text_start_token])[0][0]
        self.end_id = self.tokenizer([text_end_token])[0][0]
        self.pad_id = 0 if text_pad_token is None else self.tokenizer([text_pad_token])[0][0]
        self.text_max_length = int(text_max_length)
        self.device = torch.device(device)

    def forward(self, text: Union[List[str], str]) -> Tensor:
        single = isinstance(text, str)
        L = self.text_max_length
        rows = [[self.start_id] + ids[:L - 2] + [self.end_id] for ids in self.tokenizer.encode_batch([text] if single else list(text))]
        longest = max(len(r) for r in rows)
        # ToTensor(padding_value=0) pads to the longest row of the batch with 0, PadTransform then pads to L with pad_id
        out = torch.full((len(rows), max(L, longest)), self.pad_id, dtype=torch.long)
        for i, r in enumerate(rows):
            out[i, :len(r)] = torch.tensor(r, dtype=torch.long)
            out[i, len(r):longest] = 0
        out = out[0] if single else out
        return out.to(self.device, non_blocking=True) if self.device.type != "cpu" else out


class CLIPTransform(nn.Module):
    """clip_transform.py:355-420: `(image, text) -> (image tensor, token ids)`."""

    def __init__(self, image_size: Union[int, Tuple[int, int]] = 224, image_interpolation="bicubic",
                 image_mean: Tuple[float, float, float] = CLIP_DEFAULT_MEAN,
                 image_std: Tuple[float, float, float] = CLIP_DEFAULT_STD, text_max_length: int = 77, is_train: bool = True,
                 text_start_token: str = "<|startoftext|>", text_end_token: str = "<|endoftext|>",
                 text_pad_token: Optional[str] = None, text_bpe_merges_path: str = CLIP_DEFAULT_VOCAB_BPE_PATH,
                 num_merges: Optional[int] = 48894, device: Union[str, torch.device] = "cuda") -> None:
        super().__init__()
        self.image_transform = CLIPImageTransform(image_size, image_interpolation, image_mean, image_std, is_train, device)
        self.text_transform = CLIPTextTransform(text_max_length, text_start_token, text_end_token, text_pad_token,
                                                text_bpe_merges_path, num_merges, device)

    def forward(self, image, text: Union[List[str], str]) -> Tuple[Tensor, Tensor]:
        return self.image_transform(image), self.text_transform(text)
