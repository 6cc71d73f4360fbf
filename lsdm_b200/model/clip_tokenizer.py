"""Host-side BPE tokeniser of the CLIP text tower -- what the reference calls as ``clip.tokenize(raw_text, context_length=22,
truncate=True)`` (model/sdm.py:245-256).  The ``clip`` package (openai/CLIP, an unpinned git dependency of the reference,
README.md:27) is not part of /root/reference and not installed; this is a restatement of its published algorithm
(``clip/simple_tokenizer.py``, ``clip/clip.py::tokenize``):

* text: optional ftfy fix (if the package is present), double ``html.unescape``, whitespace collapsed, lower-cased;
* split by the CLIP pattern (special tokens, English contractions, letter runs, single digits, other non-space runs);
* each piece -> UTF-8 bytes -> the printable byte alphabet, last symbol tagged ``</w>``, then byte-pair merges applied greedily
  in the rank order of the vocabulary file until no ranked pair is left;
* ids: 256 byte symbols, the same 256 with ``</w>``, one id per merge (file order), ``<|startoftext|>``, ``<|endoftext|>``;
* ``tokenize``: ``[sot] + ids + [eot]`` per text, zero-padded to ``context_length``; longer inputs raise unless ``truncate`` (then
  the last kept position becomes ``eot``).

The vocabulary (``bpe_simple_vocab_16e6.txt.gz``, 1.3 MB, shipped inside the ``clip`` package) is NOT in this repository and not
in the build image: pass its path (``ClipBpeTokenizer(path)``, ``SceneDiffusionModel.set_tokenizer_vocab(path)`` or the
``LSDM_CLIP_BPE`` environment variable).  Parity: pinned on CPU against ``transformers.CLIPTokenizer`` -- an independent
implementation of the same algorithm -- on a synthetic merge table (tests/test_host_logic.py); unpinned against the ``clip``
package itself (absent).
"""
from __future__ import annotations

import gzip
import html
import os

import regex

_SOT, _EOT = "<|startoftext|>", "<|endoftext|>"
_SPLIT = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)


def _byte_alphabet():
    """byte value -> printable unicode character: the 188 printable latin-1 bytes map to themselves, the other 68 to U+0100..."""
    keep = [b for b in range(256) if 33 <= b <= 126 or 161 <= b <= 172 or 174 <= b <= 255]
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    # id order of the vocabulary: the kept bytes first (ascending), then the remapped ones in byte order
    order = keep + [b for b in range(256) if b not in keep]
    return table, [table[b] for b in order]


class ClipBpeTokenizer:
    def __init__(self, vocab_path=None, merges=None, n_merges=49152 - 256 - 2):
        """``vocab_path``: the gzip'd (or plain) merge list of openai/CLIP (first line is a header); or ``merges``: a list of
        ``(left, right)`` pairs in rank order."""
        if merges is None:
            vocab_path = vocab_path or os.environ.get("LSDM_CLIP_BPE")
            if not vocab_path or not os.path.exists(vocab_path):
                raise FileNotFoundError("CLIP BPE vocabulary not found: pass the path of clip's bpe_simple_vocab_16e6.txt.gz "
                                        "(argument, SceneDiffusionModel.set_tokenizer_vocab, or LSDM_CLIP_BPE)")
            opener = gzip.open if vocab_path.endswith(".gz") else open
            with opener(vocab_path, "rb") as f:
                lines = f.read().decode("utf-8").split("\n")
            merges = [tuple(l.split()) for l in lines[1:1 + n_merges]]
            merges = [m for m in merges if len(m) == 2]
        self.byte_sym, alphabet = _byte_alphabet()
        symbols = alphabet + [s + "</w>" for s in alphabet] + [a + b for a, b in merges] + [_SOT, _EOT]
        self.ids = {s: i for i, s in enumerate(symbols)}
        self.rank = {tuple(m): i for i, m in enumerate(merges)}
        self.sot, self.eot = self.ids[_SOT], self.ids[_EOT]
        self._memo = {}
        try:
            import ftfy
            self._fix = ftfy.fix_text
        except Exception:  # ftfy is an optional dependency of clip; without it mojibake is left as it is
            self._fix = lambda s: s

    def _merge_word(self, piece):
        """Greedy lowest-rank-first pair merging of one pre-token given as a string over the byte alphabet."""
        got = self._memo.get(piece)
        if got is not None:
            return got
        parts = list(piece[:-1]) + [piece[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = None, None
            for pair in zip(parts, parts[1:]):
                r = self.rank.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            merged, i = [], 0
            while i < len(parts):
                if i + 1 < len(parts) and (parts[i], parts[i + 1]) == best:
                    merged.append(parts[i] + parts[i + 1])
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        self._memo[piece] = parts
        return parts

    def encode(self, text):
        text = html.unescape(html.unescape(self._fix(text))).strip()
        text = regex.sub(r"\s+", " ", text).strip().lower()
        out = []
        for piece in _SPLIT.findall(text):
            if piece in (_SOT, _EOT):
                out.append(self.ids[piece])
                continue
            sym = "".join(self.byte_sym[b] for b in piece.encode("utf-8"))
            out.extend(self.ids[p] for p in self._merge_word(sym))
        return out

    def tokenize(self, texts, context_length=77, truncate=False):
        """-> int64 tensor [len(texts), context_length] (the layout ``lsdm_clip_encode_text`` / ``ClipTextTower`` take)."""
        import torch

        if isinstance(texts, str):
            texts = [texts]
        res = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [self.sot] + self.encode(t) + [self.eot]
            if len(ids) > context_length:
                if not truncate:
                    raise RuntimeError(f"Input {t} is too long for context length {context_length}")
                ids = ids[:context_length]
                ids[-1] = self.eot
            res[i, :len(ids)] = torch.tensor(ids)
        return res

    __call__ = tokenize
