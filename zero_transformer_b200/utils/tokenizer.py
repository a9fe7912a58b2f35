"""Byte-level BPE tokenizer (GPT-2 / GPT-NeoX family) with no runtime dependency on ``transformers``.

The reference app downloads ``EleutherAI/gpt-neox-20b`` through ``GPTNeoXTokenizerFast`` (``/root/reference/app.py:27``)
and the training shards are tokenised with the same vocabulary (50 277 entries, padded to 50 304 in the model table).
This module reads that tokenizer's own file format -- a Hugging Face ``tokenizer.json`` (``model.vocab``, ``model.merges``,
``added_tokens``, byte-level pre-tokeniser) or the older ``vocab.json`` + ``merges.txt`` pair -- and implements the same
algorithm: split with the GPT-2 regex, map bytes to the printable-unicode alphabet, merge pairs by rank.  It can also
*train* a byte-level BPE on a local corpus (``train_bpe``), so a deployment without network access can build the
tokenizer that its shards were (or will be) encoded with, and ``save`` writes a ``tokenizer.json`` that both this class
and ``tokenizers``/``transformers`` load.

``load_tokenizer(path)`` resolution order: explicit path (file or directory) -> ``$ZTB_TOKENIZER`` -> a locally cached
``gpt-neox-20b`` snapshot under ``$HF_HOME`` -> ``ByteTokenizer`` (UTF-8 bytes; announced loudly, because its ids mean
nothing to a model trained on BPE ids).
"""
from __future__ import annotations

import glob
import json
import os
import sys
from collections import Counter
from functools import lru_cache
from typing import Dict, Iterable, List, Optional, Tuple

try:  # the GPT-2 split pattern needs \p{L} / \p{N}
    import regex as _re
except ImportError:  # pragma: no cover - regex is part of the image
    _re = None

GPT2_SPLIT = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """The reversible byte -> printable character table of GPT-2's byte-level BPE."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


class ByteTokenizer:
    """Last-resort tokenizer (UTF-8 bytes) when no BPE vocabulary can be found offline."""

    eos_token_id = 0
    vocab_size = 256

    def encode(self, s: str) -> List[int]:
        return list(s.encode("utf-8"))

    def decode(self, ids) -> str:
        return bytes(int(i) % 256 for i in ids).decode("utf-8", errors="replace")


class BPETokenizer:
    def __init__(self, vocab: Dict[str, int], merges: List[Tuple[str, str]], special: Optional[Dict[str, int]] = None,
                 eos_token: str = "<|endoftext|>", add_prefix_space: bool = False, split_pattern: str = GPT2_SPLIT):
        if _re is None:  # pragma: no cover
            raise RuntimeError("the 'regex' package is required for the byte-level BPE tokenizer")
        self.vocab = dict(vocab)
        self.special = dict(special or {})
        for tok, i in self.special.items():
            self.vocab.setdefault(tok, i)
        self.inv = {i: t for t, i in self.vocab.items()}
        self.ranks = {tuple(m): r for r, m in enumerate(merges)}
        self.merges = [tuple(m) for m in merges]
        self.b2u = bytes_to_unicode()
        self.u2b = {u: b for b, u in self.b2u.items()}
        self.pat = _re.compile(split_pattern)
        self.add_prefix_space = add_prefix_space
        self.eos_token = eos_token
        self.eos_token_id = self.vocab.get(eos_token)
        self._special_pat = _re.compile("|".join(_re.escape(t) for t in sorted(self.special, key=len, reverse=True))) if self.special else None
        self._cache: Dict[str, List[str]] = {}

    # ------------------------------------------------------------------ properties
    @property
    def vocab_size(self) -> int:
        return max(self.inv) + 1 if self.inv else 0

    def __len__(self) -> int:
        return self.vocab_size

    # ------------------------------------------------------------------ core
    def _bpe(self, word: str) -> List[str]:
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        parts = list(word)
        while len(parts) > 1:
            best, best_rank = -1, None
            for i in range(len(parts) - 1):
                r = self.ranks.get((parts[i], parts[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = parts[best], parts[best + 1]
            out, i = [], 0
            while i < len(parts):   # merge every occurrence of the winning pair, left to right
                if i < len(parts) - 1 and parts[i] == a and parts[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(parts[i])
                    i += 1
            parts = out
        if len(self._cache) < 200000:
            self._cache[word] = parts
        return parts

    def _encode_ordinary(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in self.pat.findall(text):
            mapped = "".join(self.b2u[b] for b in piece.encode("utf-8"))
            for tok in self._bpe(mapped):
                i = self.vocab.get(tok)
                if i is None:   # cannot happen with a complete byte alphabet; degrade to per-character ids
                    ids.extend(self.vocab[c] for c in tok if c in self.vocab)
                else:
                    ids.append(i)
        return ids

    def encode(self, text: str, allow_special: bool = True) -> List[int]:
        if self.add_prefix_space and text and not text[0].isspace():
            text = " " + text
        if not (allow_special and self._special_pat):
            return self._encode_ordinary(text)
        ids: List[int] = []
        pos = 0
        for m in self._special_pat.finditer(text):
            ids.extend(self._encode_ordinary(text[pos:m.start()]))
            ids.append(self.special[m.group(0)])
            pos = m.end()
        ids.extend(self._encode_ordinary(text[pos:]))
        return ids

    def decode(self, ids: Iterable[int], skip_special: bool = False) -> str:
        buf = bytearray()
        out: List[str] = []

        def flush():
            if buf:
                out.append(bytes(buf).decode("utf-8", errors="replace"))
                buf.clear()

        special_ids = set(self.special.values())
        for i in ids:
            i = int(i)
            tok = self.inv.get(i)
            if tok is None:
                continue   # padded vocabulary rows (50 277 .. 50 303) have no string
            if i in special_ids:
                flush()
                if not skip_special:
                    out.append(tok)
                continue
            for ch in tok:
                b = self.u2b.get(ch)
                if b is None:
                    flush()
                    out.append(ch)
                else:
                    buf.append(b)
        flush()
        return "".join(out)

    # ------------------------------------------------------------------ files
    @classmethod
    def from_file(cls, path: str) -> "BPETokenizer":
        """``tokenizer.json`` (HF tokenizers), a directory containing one, or a directory with ``vocab.json`` + ``merges.txt``."""
        if os.path.isdir(path):
            tj = os.path.join(path, "tokenizer.json")
            if os.path.exists(tj):
                return cls.from_file(tj)
            vj, mt = os.path.join(path, "vocab.json"), os.path.join(path, "merges.txt")
            if os.path.exists(vj) and os.path.exists(mt):
                vocab = json.load(open(vj, encoding="utf-8"))
                merges = []
                for line in open(mt, encoding="utf-8"):
                    line = line.rstrip("\n")
                    if line and not line.startswith("#version"):
                        a, b = line.split(" ")
                        merges.append((a, b))
                return cls(vocab, merges, {"<|endoftext|>": vocab["<|endoftext|>"]} if "<|endoftext|>" in vocab else None)
            raise FileNotFoundError(f"{path}: no tokenizer.json and no vocab.json + merges.txt")
        d = json.load(open(path, encoding="utf-8"))
        model = d["model"]
        if model.get("type", "BPE") != "BPE":
            raise ValueError(f"{path}: only byte-level BPE tokenizers are supported (got {model.get('type')})")
        merges = [tuple(m.split(" ")) if isinstance(m, str) else tuple(m) for m in model["merges"]]
        special = {t["content"]: int(t["id"]) for t in d.get("added_tokens", [])}
        pre = d.get("pre_tokenizer") or {}
        pres = pre.get("pretokenizers", [pre]) if pre else []
        prefix = any(p.get("type") == "ByteLevel" and p.get("add_prefix_space") for p in pres)
        eos = "<|endoftext|>" if "<|endoftext|>" in special or "<|endoftext|>" in model["vocab"] else next(iter(special), "<|endoftext|>")
        return cls(model["vocab"], merges, special, eos_token=eos, add_prefix_space=prefix)

    def save(self, path: str) -> None:
        """Write a ``tokenizer.json`` that this class, ``tokenizers.Tokenizer.from_file`` and ``transformers`` can load."""
        base_vocab = dict(self.vocab)   # special tokens stay in the model vocabulary, as the HF trainers write them
        d = {
            "version": "1.0", "truncation": None, "padding": None,
            "added_tokens": [{"id": i, "content": t, "single_word": False, "lstrip": False, "rstrip": False, "normalized": False,
                              "special": True} for t, i in sorted(self.special.items(), key=lambda kv: kv[1])],
            "normalizer": None,
            "pre_tokenizer": {"type": "ByteLevel", "add_prefix_space": self.add_prefix_space, "trim_offsets": True, "use_regex": True},
            "post_processor": {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": True},
            "decoder": {"type": "ByteLevel", "add_prefix_space": False, "trim_offsets": True, "use_regex": True},
            "model": {"type": "BPE", "dropout": None, "unk_token": None, "continuing_subword_prefix": None, "end_of_word_suffix": None,
                      "fuse_unk": False, "byte_fallback": False, "ignore_merges": False, "vocab": base_vocab,
                      "merges": [[a, b] for a, b in self.merges]},
        }
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w", encoding="utf-8") as f:
            json.dump(d, f, ensure_ascii=False)


def train_bpe(texts: Iterable[str], vocab_size: int, special_tokens: Tuple[str, ...] = ("<|endoftext|>",), min_frequency: int = 2) -> BPETokenizer:
    """Train a byte-level BPE on an in-memory / streamed corpus (word-frequency formulation: split with the GPT-2 regex,
    count distinct words, then merge the most frequent adjacent pair until the vocabulary is full).  Meant for building a
    tokenizer offline from a few hundred MB of local text; for the published GPT-NeoX vocabulary load its tokenizer.json."""
    b2u = bytes_to_unicode()
    pat = _re.compile(GPT2_SPLIT)
    words: Counter = Counter()
    for t in texts:
        for piece in pat.findall(t):
            words["".join(b2u[b] for b in piece.encode("utf-8"))] += 1
    alphabet = sorted(b2u.values())
    vocab: Dict[str, int] = {}
    for tok in special_tokens:
        vocab[tok] = len(vocab)
    n_special = len(vocab)
    for ch in alphabet:
        vocab[ch] = len(vocab)
    splits = {w: list(w) for w in words}
    merges: List[Tuple[str, str]] = []
    while len(vocab) < vocab_size:
        pairs: Counter = Counter()
        for w, parts in splits.items():
            f = words[w]
            for a, b in zip(parts, parts[1:]):
                pairs[(a, b)] += f
        if not pairs:
            break
        # deterministic tie-break: highest count, then lexicographically smallest pair
        (a, b), cnt = min(pairs.items(), key=lambda kv: (-kv[1], kv[0]))
        if cnt < min_frequency:
            break
        merges.append((a, b))
        vocab[a + b] = len(vocab)
        for w, parts in splits.items():
            if len(parts) < 2:
                continue
            out, i = [], 0
            while i < len(parts):
                if i < len(parts) - 1 and parts[i] == a and parts[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(parts[i])
                    i += 1
            splits[w] = out
    special = {t: i for t, i in vocab.items() if i < n_special}
    return BPETokenizer(vocab, merges, special, eos_token=special_tokens[0] if special_tokens else "<|endoftext|>")


def _cached_neox() -> Optional[str]:
    roots = [os.environ.get("HF_HOME"), os.path.expanduser("~/.cache/huggingface")]
    for r in roots:
        if not r:
            continue
        hits = glob.glob(os.path.join(r, "hub", "models--EleutherAI--gpt-neox-20b", "snapshots", "*", "tokenizer.json"))
        if hits:
            return sorted(hits)[-1]
    return None


def load_tokenizer(path: Optional[str] = None, quiet: bool = False):
    """See the module docstring for the resolution order."""
    cand = path or os.environ.get("ZTB_TOKENIZER") or _cached_neox()
    if cand:
        return BPETokenizer.from_file(cand)
    if not quiet:
        print("[tokenizer] no BPE vocabulary found (pass --tokenizer <tokenizer.json|dir>, set ZTB_TOKENIZER, or place the "
              "gpt-neox-20b snapshot in the HF cache): falling back to raw UTF-8 bytes -- ids will NOT match a model trained on "
              "GPT-NeoX token ids", file=sys.stderr)
    return ByteTokenizer()
