"""Warm-start depth extension (Gopher, appendix G3.3): grow an ``n``-block
parameter tree into a ``k*n``-block one by giving source block ``i`` to target
blocks ``k*i .. k*i+k-1``.  The reference hard-codes 18 -> 36
(``src/utils/extend_params.py:12-49``); here any integer multiple works and the
same function extends Adam moments.
"""
from __future__ import annotations

import re
from typing import Dict

import torch

_BLOCK = re.compile(r"^blocks\.(\d+)\.(.+)$")


def create_mapping(n_src: int, n_dst: int) -> Dict[int, int]:
    """target block index -> source block index."""
    if n_dst % n_src != 0:
        raise ValueError(f"target depth {n_dst} is not a multiple of source depth {n_src}")
    k = n_dst // n_src
    return {j: j // k for j in range(n_dst)}


def extend_params(params: Dict[str, torch.Tensor], n_dst: int) -> Dict[str, torch.Tensor]:
    src_blocks = sorted({int(m.group(1)) for m in map(_BLOCK.match, params) if m})
    n_src = len(src_blocks)
    if n_src == 0:
        raise ValueError("no 'blocks.<i>.*' entries in the source parameters")
    mapping = create_mapping(n_src, n_dst)
    per_block: Dict[str, Dict[int, torch.Tensor]] = {}
    out: Dict[str, torch.Tensor] = {}
    for name, t in params.items():
        m = _BLOCK.match(name)
        if m:
            per_block.setdefault(m.group(2), {})[int(m.group(1))] = t
        else:
            out[name] = t.clone()
    for suffix, by_idx in per_block.items():
        for j in range(n_dst):
            out[f"blocks.{j}.{suffix}"] = by_idx[mapping[j]].clone()
    return out
