"""Flat parameter layout and ZeRO-1 shard map.

Capability parity: the reference shards every optimizer-state tensor over the
single ``"dp"`` mesh axis by per-tensor regex rules (``src/partitioning/partition.py:49-140``)
and asserts divisibility.  Here every parameter is a *region* of one flat buffer;
region ``i`` is split into ``world`` equal contiguous chunks and rank ``r`` owns
chunk ``r`` (for a 2-D ``(out, in)`` weight that is a block of rows, the same
thing ``P("dp", None)`` expresses).  Regions are padded so chunks are 16-byte
aligned, which removes the divisibility requirement.  Regions are grouped into
*buckets* (one per transformer block, one for the embedding + final norm) -- the
unit at which the reduce-scatter+Adam kernel is launched during backward and at
which the all-gather is consumed during forward.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch

ALIGN_ELEMS = 64  # chunk granularity in elements (128 B of bf16 = one TMA swizzle row)


@dataclass
class Region:
    name: str
    shape: Tuple[int, ...]
    numel: int
    offset: int        # element offset of the region in the flat buffers
    padded: int        # region length in elements (multiple of world*ALIGN)
    chunk: int         # padded // world
    decay: bool        # receives weight decay
    bucket: int
    shard_offset: int = 0  # offset of this region's chunk inside a rank's flat shard

    def chunk_range(self, rank: int) -> Tuple[int, int]:
        s = self.offset + rank * self.chunk
        return s, s + self.chunk


@dataclass
class Bucket:
    index: int
    name: str
    regions: List[Region] = field(default_factory=list)

    @property
    def start(self) -> int:
        return self.regions[0].offset

    @property
    def end(self) -> int:
        r = self.regions[-1]
        return r.offset + r.padded

    @property
    def shard_start(self) -> int:
        return self.regions[0].shard_offset

    @property
    def shard_end(self) -> int:
        r = self.regions[-1]
        return r.shard_offset + r.chunk


def default_bucket_of(name: str) -> str:
    """``blocks.<i>.*`` -> one bucket per block; everything else -> ``embed``."""
    parts = name.split(".")
    if len(parts) > 2 and parts[0] == "blocks" and parts[1].isdigit() and not parts[2].startswith("ln"):
        return f"block{int(parts[1])}"
    # embedding, final norm and every LayerNorm scale: consumed before the block's first GEMM, so they
    # travel with the standalone all-gather at the start of the step (the per-block buckets are
    # gathered inside their QKV GEMM)
    return "embed"


class FlatLayout:
    def __init__(self, named_shapes: Sequence[Tuple[str, Tuple[int, ...]]], decay: Dict[str, bool],
                 world: int, bucket_of=default_bucket_of, align: int = ALIGN_ELEMS):
        self.world = int(world)
        self.align = int(align)
        order: List[str] = []
        groups: Dict[str, List[Tuple[str, Tuple[int, ...]]]] = {}
        for name, shape in named_shapes:
            b = bucket_of(name)
            if b not in groups:
                groups[b] = []
                order.append(b)
            groups[b].append((name, tuple(shape)))
        self.buckets: List[Bucket] = []
        self.regions: List[Region] = []
        self.by_name: Dict[str, Region] = {}
        off = 0
        soff = 0
        quantum = self.world * self.align
        for bi, bname in enumerate(order):
            bucket = Bucket(bi, bname)
            for name, shape in groups[bname]:
                n = 1
                for s in shape:
                    n *= int(s)
                padded = ((n + quantum - 1) // quantum) * quantum
                reg = Region(name, shape, n, off, padded, padded // self.world, bool(decay.get(name, len(shape) >= 2)), bi, soff)
                off += padded
                soff += reg.chunk
                bucket.regions.append(reg)
                self.regions.append(reg)
                self.by_name[name] = reg
            self.buckets.append(bucket)
        self.total = off          # elements in a full flat buffer
        self.shard_total = soff   # elements in one rank's shard (= total // world)
        assert self.total == self.shard_total * self.world

    # ---- helpers -----------------------------------------------------
    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        r = self.by_name[name]
        return flat[r.offset: r.offset + r.numel].view(r.shape)

    def segment_table(self, bucket: int | None = None) -> torch.Tensor:
        """int64 ``(n, 4)`` rows ``[flat_offset, chunk, shard_offset, decay]`` consumed
        by the fused reduce-scatter+Adam and all-gather kernels."""
        regs = self.regions if bucket is None else self.buckets[bucket].regions
        return torch.tensor([[r.offset, r.chunk, r.shard_offset, int(r.decay)] for r in regs], dtype=torch.int64)

    def decay_mask_shard(self, device=None) -> torch.Tensor:
        """fp32 0/1 vector over one rank's shard (identical on every rank)."""
        m = torch.zeros(self.shard_total, dtype=torch.float32, device=device)
        for r in self.regions:
            if r.decay:
                m[r.shard_offset: r.shard_offset + r.chunk] = 1.0
        return m

    def gather_shard_from_flat(self, flat: torch.Tensor, rank: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """Copy rank ``rank``'s chunks of a full flat buffer into a shard-shaped tensor."""
        if out is None:
            out = torch.empty(self.shard_total, dtype=flat.dtype, device=flat.device)
        for r in self.regions:
            s, e = r.chunk_range(rank)
            out[r.shard_offset: r.shard_offset + r.chunk].copy_(flat[s:e])
        return out

    def scatter_shard_to_flat(self, shard: torch.Tensor, flat: torch.Tensor, rank: int) -> None:
        for r in self.regions:
            s, e = r.chunk_range(rank)
            flat[s:e].copy_(shard[r.shard_offset: r.shard_offset + r.chunk])

    def describe(self) -> str:
        lines = [f"FlatLayout world={self.world} total={self.total} shard={self.shard_total} buckets={len(self.buckets)}"]
        for b in self.buckets:
            lines.append(f"  [{b.index}] {b.name}: {len(b.regions)} regions, {b.end - b.start} elems")
        return "\n".join(lines)
