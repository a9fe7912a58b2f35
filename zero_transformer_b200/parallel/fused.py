"""Fused ZeRO-1 communication backend (NVLink peer memory + hand-written kernels).

``pack_segments`` builds the device-side segment table consumed by the
reduce-scatter+AdamW kernel (``csrc/optim.cu``).  ``FusedComm`` owns the symmetric flat
parameter / gradient buffers and drives the per-step protocol.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

ADAM_TILE = 2048  # elements per CTA tile; must match kAdamTile in csrc/optim.cu

_SEG_DTYPE = np.dtype([("flat_off", "<i8"), ("shard_off", "<i8"), ("len", "<i8"), ("tile_begin", "<i8"),
                       ("decay", "<i4"), ("pad", "<i4")])


def pack_segments(segments: Sequence[Tuple[int, int, int, bool]], device) -> Tuple[torch.Tensor, int]:
    """``[(flat_off, shard_off, len, decay)]`` -> (uint8 CUDA tensor of packed ``ZtbSegment``, total tiles)."""
    arr = np.zeros(len(segments), dtype=_SEG_DTYPE)
    tiles = 0
    for i, (fo, so, ln, dec) in enumerate(segments):
        if ln % 8 or fo % 8 or so % 4:
            raise ValueError("segment offsets/lengths must be 8-element aligned")
        arr[i] = (fo, so, ln, tiles, int(bool(dec)), 0)
        tiles += (ln + ADAM_TILE - 1) // ADAM_TILE
    t = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    return t, tiles
