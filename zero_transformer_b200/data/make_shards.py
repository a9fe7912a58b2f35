"""Build token shards + an index file for ``StreamingTokenDataset`` (the reference consumes
pre-tokenised WebDataset tars listed in ``data/index/*.index``, ``main_zero.py:368-421``, each sample an
``input_id.pth`` of ``max_context`` tokens; this is the local equivalent).

    python -m zero_transformer_b200.data.make_shards --tokens corpus.npy --out data/shards \
        --max-context 2048 --samples-per-shard 4096 --index data/index/train.index

``--tokens`` is a 1-D integer ``.npy`` (or raw uint16 ``.bin``) token stream; ``--text`` + ``--tokenizer``
tokenises a UTF-8 text file with a local HF tokenizer directory first.
"""
from __future__ import annotations

import argparse
import os

import numpy as np


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--tokens", default=None)
    p.add_argument("--text", default=None)
    p.add_argument("--tokenizer", default=None)
    p.add_argument("--out", required=True)
    p.add_argument("--index", required=True)
    p.add_argument("--max-context", type=int, default=2048)
    p.add_argument("--samples-per-shard", type=int, default=4096)
    p.add_argument("--prefix", default="shard")
    a = p.parse_args(argv)
    if a.tokens:
        toks = np.memmap(a.tokens, dtype=np.uint16, mode="r") if a.tokens.endswith(".bin") else np.load(a.tokens, mmap_mode="r")
    elif a.text:
        from transformers import AutoTokenizer

        tk = AutoTokenizer.from_pretrained(a.tokenizer, local_files_only=True)
        with open(a.text, "r", encoding="utf-8") as f:
            toks = np.asarray(tk.encode(f.read()), dtype=np.int32)
    else:
        raise SystemExit("need --tokens or --text")
    toks = np.asarray(toks).reshape(-1)
    n = toks.shape[0] // a.max_context
    rows = toks[: n * a.max_context].reshape(n, a.max_context).astype(np.int32)
    os.makedirs(a.out, exist_ok=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.index)), exist_ok=True)
    names = []
    for s in range(0, n, a.samples_per_shard):
        path = os.path.join(a.out, f"{a.prefix}-{s // a.samples_per_shard:06d}.npy")
        np.save(path, rows[s: s + a.samples_per_shard])
        names.append(os.path.relpath(path, os.path.dirname(os.path.abspath(a.index))))
    with open(a.index, "w") as f:
        f.write("# one shard per line, relative to this file\n" + "\n".join(names) + "\n")
    print(f"wrote {len(names)} shard(s), {n} samples of {a.max_context} tokens, index {a.index}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
