"""CPU restatement (pure Python) of the CLIP text transform — TEST INFRASTRUCTURE ONLY.

Follows torchmultimodal/transforms/clip_transform.py: bytes_to_unicode :31-55, CLIPBPETokenizer.__init__ :106-141 (merge
ranks, vocabulary order: 256 byte symbols, the same + "</w>", one entry per merge line, bos, eos), .bpe :147-185 (lowest
rank pair first, every occurrence merged left to right), .encode :187-197 (lower, strip, regex split, byte alphabet), and
CLIPTextTransform :244-298 with text_transforms.Truncate / AddToken / ToTensor(padding_value=0) / PadTransform.
Pinned by tests/test_clip_text_transform_cpu.py against tests/golden/clip_text_golden.pt (ids produced by the unmodified
reference classes on the synthetic merges file tests/golden/clip_bpe_merges.bpe; generator
tests/golden/make_clip_text_golden.py).
"""
from typing import Dict, List, Optional, Tuple

import regex

PATTERN = regex.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
                        regex.IGNORECASE)


def byte_alphabet() -> List[str]:
    """alphabet[b] = the printable stand-in of byte b."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, nxt = {}, 256
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(nxt)
            nxt += 1
    return [table[b] for b in keep] + [table[b] for b in range(256) if b not in keep], table


class Tokenizer:
    def __init__(self, merges_text: str, bos: str = "<|startoftext|>", eos: str = "<|endoftext|>",
                 num_merges: Optional[int] = None):
        order, self.table = byte_alphabet()
        lines = merges_text.split("\n")[1:]
        lines = lines[:(num_merges or len(lines))]
        pairs = [tuple(ln.split()) for ln in lines]
        self.rank: Dict[Tuple[str, ...], int] = {p: i for i, p in enumerate(pairs)}
        vocab = order + [s + "</w>" for s in order] + ["".join(p) for p in pairs] + [bos, eos]
        self.ids = {s: i for i, s in enumerate(vocab)}
        self.special = (bos, eos)

    def word(self, piece: str) -> List[int]:
        text = "".join(self.table[b] for b in piece.encode("utf-8"))
        if text in self.special:
            return [self.ids[text]]
        syms = list(text[:-1]) + [text[-1] + "</w>"]
        while len(syms) > 1:
            ranked = [(self.rank[(a, b)], i) for i, (a, b) in enumerate(zip(syms, syms[1:])) if (a, b) in self.rank]
            if not ranked:
                break
            _, at = min(ranked)
            a, b = syms[at], syms[at + 1]
            out, i = [], 0
            while i < len(syms):
                if i + 1 < len(syms) and syms[i] == a and syms[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(syms[i])
                    i += 1
            syms = out
        return [self.ids[s] for s in syms]

    def encode(self, text: str) -> List[int]:
        out: List[int] = []
        for piece in PATTERN.findall(text.lower().strip()):
            out.extend(self.word(piece))
        return out


def text_transform(tok: Tokenizer, texts: List[str], max_len: int = 77, pad_token: Optional[str] = None) -> List[List[int]]:
    """CLIPTextTransform.forward on a list: rows of length max(max_len, longest) as nested lists."""
    start, end = tok.encode(tok.special[0])[0], tok.encode(tok.special[1])[0]
    pad = 0 if pad_token is None else tok.encode(pad_token)[0]
    rows = [[start] + tok.encode(t)[:max_len - 2] + [end] for t in texts]
    longest = max(len(r) for r in rows)
    return [r + [0] * (longest - len(r)) + [pad] * (max(max_len, longest) - longest) for r in rows]
