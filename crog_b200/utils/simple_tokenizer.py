"""CLIP byte-pair tokenizer: the host step that turns a referring expression into the `word` ids of
``CROG.forward``.

Mirrors the reference surface ``utils/simple_tokenizer.py:62-132`` (``SimpleTokenizer(bpe_path)``,
``.encode``, ``.decode``, ``.encoder`` / ``.decoder``) and ``utils/dataset.py:57-98`` (``tokenize``): same ids
for the same text and the same merge table.  The merge table itself (``bpe_simple_vocab_16e6.txt.gz``, the file
the reference ships next to its tokenizer) is data, not code, and is NOT copied into this repository: pass its
path, or set ``CROG_BPE_PATH``, or keep the reference checkout at one of the default locations.

The algorithm is the published CLIP / GPT-2 byte-level BPE, written here over integer symbol ids:

* text -> (optional ftfy) -> HTML-unescape twice -> strip -> collapse whitespace -> lower case
  (utils/simple_tokenizer.py:50-59,121-124); ``ftfy`` is optional - without it already-clean text is unchanged,
* split with the CLIP pattern (specials, English contractions, letter runs, single digits, other runs),
* each UTF-8 byte becomes one printable code point (the 256-entry GPT-2 byte alphabet), the last symbol of a
  word carries the ``</w>`` marker,
* adjacent symbols are merged lowest-rank-first until no ranked pair is left (ranks = line number in the merge
  file, first 48 894 merges), every resulting symbol is a vocabulary id.

Vocabulary layout (utils/simple_tokenizer.py:69-75): 256 byte symbols, the same 256 with ``</w>``, one entry per
merge, ``<|startoftext|>`` = 49406, ``<|endoftext|>`` = 49407.
"""
from __future__ import annotations

import gzip
import html
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import regex

try:  # optional: mojibake repair exactly as the reference does when it is installed
    import ftfy as _ftfy
except Exception:  # pragma: no cover - not installed in this image
    _ftfy = None

N_MERGES = 49152 - 256 - 2
SOT, EOT = "<|startoftext|>", "<|endoftext|>"
END = "</w>"
_PATTERN = regex.compile(
    r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)
_SPACE = regex.compile(r"\s+")
_DEFAULT_LOCATIONS = (
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpe_simple_vocab_16e6.txt.gz"),
    "/root/reference/utils/bpe_simple_vocab_16e6.txt.gz",
)


def default_bpe() -> str:
    """Path of the merge table: $CROG_BPE_PATH, a copy next to this file, or the reference checkout."""
    env = os.environ.get("CROG_BPE_PATH")
    if env:
        return env
    for p in _DEFAULT_LOCATIONS:
        if os.path.exists(p):
            return p
    raise FileNotFoundError(
        "CLIP merge table bpe_simple_vocab_16e6.txt.gz not found: pass bpe_path=..., set CROG_BPE_PATH, or place the "
        "file next to crog_b200/utils/simple_tokenizer.py (it ships with the reference under utils/)")


def byte_alphabet() -> List[str]:
    """The GPT-2 byte -> printable code point table as a list indexed by byte value: printable Latin-1 bytes map to
    themselves, the 68 others to U+0100.. in increasing byte order."""
    keep = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, spare = [], 0
    for b in range(256):
        if b in keep:
            table.append(chr(b))
        else:
            table.append(chr(256 + spare))
            spare += 1
    return table


def vocabulary_order(alphabet: Sequence[str]) -> List[str]:
    """Symbols 0..255 of the vocabulary are the byte symbols ordered as the reference lists them: the three printable
    ranges first, then the remapped bytes (utils/simple_tokenizer.py:16-35)."""
    head = list(range(0x21, 0x7F)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    seen = set(head)
    tail = [b for b in range(256) if b not in seen]
    return [alphabet[b] for b in head + tail]


def basic_clean(text: str) -> str:
    if _ftfy is not None:
        text = _ftfy.fix_text(text)
    return html.unescape(html.unescape(text)).strip()


def whitespace_clean(text: str) -> str:
    return _SPACE.sub(" ", text).strip()


class SimpleTokenizer:
    def __init__(self, bpe_path: Optional[str] = None):
        path = bpe_path or default_bpe()
        with gzip.open(path, "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        pairs = [tuple(ln.split()) for ln in lines[1:1 + N_MERGES]]  # line 0 is the version header
        if len(pairs) != N_MERGES or any(len(p) != 2 for p in pairs):
            raise ValueError(f"{path}: expected {N_MERGES} two-symbol merges after the header line")
        self._alphabet = byte_alphabet()
        base = vocabulary_order(self._alphabet)
        symbols = base + [s + END for s in base] + [a + b for a, b in pairs] + [SOT, EOT]
        self.encoder: Dict[str, int] = {s: i for i, s in enumerate(symbols)}
        self.decoder: Dict[int, str] = {i: s for s, i in self.encoder.items()}
        # merge table over ids: (left id, right id) -> (rank, merged id)
        enc = self.encoder
        self._merge: Dict[Tuple[int, int], Tuple[int, int]] = {}
        for rank, (a, b) in enumerate(pairs):
            self._merge.setdefault((enc[a], enc[b]), (rank, enc[a + b]))
        self.bpe_ranks = {p: r for r, p in enumerate(pairs)}  # reference attribute (symbol pairs -> rank)
        self._byte_id = [enc[self._alphabet[b]] for b in range(256)]
        self._byte_end_id = [enc[self._alphabet[b] + END] for b in range(256)]
        self._unbyte = {c: b for b, c in enumerate(self._alphabet)}
        self._cache: Dict[str, Tuple[int, ...]] = {SOT: (enc[SOT],), EOT: (enc[EOT],)}
        self.sot_token, self.eot_token = enc[SOT], enc[EOT]

    # ------------------------------------------------------------------ BPE over ids
    def _merge_word(self, piece: str) -> Tuple[int, ...]:
        hit = self._cache.get(piece)
        if hit is not None:
            return hit
        raw = piece.encode("utf-8")
        ids = [self._byte_id[b] for b in raw[:-1]] + [self._byte_end_id[raw[-1]]]
        merge = self._merge
        while len(ids) > 1:
            # lowest-ranked adjacent pair; all its occurrences are merged left to right in one sweep, as the
            # published algorithm does (utils/simple_tokenizer.py:80-119)
            best = None
            for i in range(len(ids) - 1):
                m = merge.get((ids[i], ids[i + 1]))
                if m is not None and (best is None or m[0] < best[0]):
                    best = (m[0], m[1], ids[i], ids[i + 1])
            if best is None:
                break
            _, new_id, left, right = best
            out, i, n = [], 0, len(ids)
            while i < n:
                if i + 1 < n and ids[i] == left and ids[i + 1] == right:
                    out.append(new_id)
                    i += 2
                else:
                    out.append(ids[i])
                    i += 1
            ids = out
        res = tuple(ids)
        self._cache[piece] = res
        return res

    def bpe(self, token: str) -> str:
        """Reference-style view of the merge result: the symbols separated by spaces."""
        return " ".join(self.decoder[i] for i in self._merge_word(token))

    # ------------------------------------------------------------------ public surface
    def encode(self, text: str) -> List[int]:
        text = whitespace_clean(basic_clean(text)).lower()
        out: List[int] = []
        for piece in _PATTERN.findall(text):
            out.extend(self._merge_word(piece))
        return out

    def decode(self, tokens: Iterable[int]) -> str:
        chars = "".join(self.decoder[int(t)] for t in tokens)
        # every symbol character (the end-of-word marker's included) is a member of the byte alphabet
        return bytearray(self._unbyte[c] for c in chars).decode("utf-8", errors="replace").replace(END, " ")


_default_tokenizer: Optional[SimpleTokenizer] = None


def get_tokenizer() -> SimpleTokenizer:
    global _default_tokenizer
    if _default_tokenizer is None:
        _default_tokenizer = SimpleTokenizer()
    return _default_tokenizer


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False, tokenizer: Optional[SimpleTokenizer] = None):
    """utils/dataset.py:57-98: [SOT] + ids + [EOT], zero padded to ``context_length`` (int64 tensor
    [len(texts), context_length]); too long inputs raise unless ``truncate`` (then the last kept id becomes EOT)."""
    import torch

    tk = tokenizer or get_tokenizer()
    if isinstance(texts, str):
        texts = [texts]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        ids = [tk.sot_token] + tk.encode(text) + [tk.eot_token]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = tk.eot_token
        result[i, :len(ids)] = torch.tensor(ids)
    return result
