"""Fused ZeRO-1 communication backend: NVLink peer memory + hand-written sm_100a kernels,
no NCCL call on the gradient reduce-scatter or the parameter all-gather.

Replaces the reference's ``pmean`` all-reduce + local slice + XLA-inserted fp32 all-gather
(``xmap_train_functions.py:83-84``, ``main_zero.py:452-460,495-500``) with

  backward ──► [signal GRAD_READY]
  reduce-scatter ⊕ scale ⊕ clip ⊕ AdamW ⊕ bf16 cast      (csrc/optim.cu; pulls every peer's
                                                           gradient chunk over NVLink, never
                                                           materialises the reduced gradient)
            ──► [signal SHARD_READY]
  all-gather (P2P pull of the updated bf16 chunks)         (csrc/comm.cu)

The flat bf16 parameter and gradient buffers and the flag pads live in symmetric memory
(``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into every peer), so
kernels address peers with ordinary pointers.  With world size 1 the same kernels run with a
single "peer" and ordinary allocations.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

ADAM_TILE = 2048  # elements per CTA tile; must match kAdamTile in csrc/optim.cu
COPY_TILE = 8192  # must match kCopyTile in csrc/comm.cu
FLAG_GRAD_READY, FLAG_SHARD_READY = 0, 1
MAX_RANKS = 8

_SEG_DTYPE = np.dtype([("flat_off", "<i8"), ("shard_off", "<i8"), ("len", "<i8"), ("tile_begin", "<i8"),
                       ("decay", "<i4"), ("pad", "<i4")])


def pack_segments(segments: Sequence[Tuple[int, int, int, int]], device, tile: int = ADAM_TILE) -> Tuple[torch.Tensor, int]:
    """``[(flat_off, shard_off, len, tag)]`` -> (uint8 CUDA tensor of packed ``ZtbSegment``, total tiles).
    ``tag`` is the weight-decay flag for the AdamW kernel and the source rank for the all-gather."""
    arr = np.zeros(max(len(segments), 1), dtype=_SEG_DTYPE)
    tiles = 0
    for i, (fo, so, ln, tag) in enumerate(segments):
        if ln % 8 or fo % 8 or so % 4:
            raise ValueError("segment offsets/lengths must be 8-element aligned")
        arr[i] = (fo, so, ln, tiles, int(tag), 0)
        tiles += (ln + tile - 1) // tile
    t = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    return t, tiles


class SymmetricBuffer:
    """A tensor allocated in symmetric memory plus the device addresses of every rank's copy."""

    def __init__(self, numel: int, dtype: torch.dtype, device, group, world: int):
        self.world = world
        if world == 1:
            self.tensor = torch.zeros(numel, dtype=dtype, device=device)
            self.ptrs = [self.tensor.data_ptr()]
            self.handle = None
            return
        import torch.distributed._symmetric_memory as symm_mem

        gname = (group or dist.group.WORLD).group_name
        try:
            symm_mem.enable_symm_mem_for_group(gname)
        except Exception:
            pass
        self.tensor = symm_mem.empty(numel, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, gname)
        self.tensor.zero_()
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)


class FusedComm:
    def __init__(self, layout, rank: int, world: int, group, device, compute_dtype, grad_dtype):
        from ..ops._build import load

        if world > MAX_RANKS:
            raise ValueError(f"fused backend supports up to {MAX_RANKS} ranks per NVSwitch domain")
        if compute_dtype != torch.bfloat16:
            raise ValueError("fused backend keeps bf16 working parameters")
        if grad_dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("grad_dtype must be bfloat16 or float32")
        self._C = load()
        self.layout, self.rank, self.world, self.group, self.device = layout, rank, world, group, device
        self.grad_fp32 = grad_dtype == torch.float32
        self._param = SymmetricBuffer(layout.total, torch.bfloat16, device, group, world)
        self._grad = SymmetricBuffer(layout.total, grad_dtype, device, group, world)
        self._flags = SymmetricBuffer(2 * MAX_RANKS, torch.int32, device, group, world)
        self.flat_param, self.flat_grad = self._param.tensor, self._grad.tensor
        # my chunk of every region (AdamW) and everybody else's chunks (all-gather)
        segs = [(r.offset + rank * r.chunk, r.shard_offset, r.chunk, int(r.decay)) for r in layout.regions]
        self.adam_segs, self.adam_tiles = pack_segments(segs, device, ADAM_TILE)
        self.n_adam = len(segs)
        pull = [(r.offset + p * r.chunk, 0, r.chunk, p) for p in range(world) if p != rank for r in layout.regions]
        self.pull_segs, self.pull_tiles = pack_segments(pull, device, COPY_TILE)
        self.n_pull = len(pull)
        self.epoch = 0
        if world > 1:
            dist.barrier(group)

    # device address of this rank's flag row for `kind`
    def _my_flags(self, kind: int) -> int:
        return self._flags.ptrs[self.rank] + 4 * kind * MAX_RANKS

    def _signal(self, kind: int, value: int) -> None:
        # slot (kind, my rank) in every rank's pad (mine included)
        self._C.signal(self._flags.ptrs, kind * MAX_RANKS + self.rank, value)

    def step(self, *, master, m, v, lr, beta1, beta2, eps, wd, step, clip, grad_scale) -> None:
        self.epoch += 1
        e = self.epoch
        if self.world > 1:
            self._signal(FLAG_GRAD_READY, e)
            grad_ptrs = self._grad.ptrs[self.rank:] + self._grad.ptrs[: self.rank]  # start with the local copy
            wait_ptr, wait_val = self._my_flags(FLAG_GRAD_READY), e
        else:
            grad_ptrs, wait_ptr, wait_val = self._grad.ptrs, 0, 0
        self._C.adamw_shard(master, m, v, self.flat_param, grad_ptrs, self.grad_fp32, self.adam_segs, self.n_adam,
                            self.adam_tiles, lr, beta1, beta2, eps, wd, clip, grad_scale, step, wait_ptr, wait_val)
        if self.world > 1:
            self._signal(FLAG_SHARD_READY, e)
            self._C.allgather_pull(self.flat_param, self._param.ptrs, self.pull_segs, self.n_pull, self.pull_tiles,
                                   self._my_flags(FLAG_SHARD_READY), e)

    def all_gather_now(self) -> None:
        """Out-of-band all-gather (checkpoint load): publish + pull under a fresh epoch."""
        if self.world == 1:
            return
        self.epoch += 1
        e = self.epoch
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        self._signal(FLAG_GRAD_READY, e)  # keep both counters in lock-step
        self._signal(FLAG_SHARD_READY, e)
        self._C.allgather_pull(self.flat_param, self._param.ptrs, self.pull_segs, self.n_pull, self.pull_tiles,
                               self._my_flags(FLAG_SHARD_READY), e)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)


def p2p_probe(env) -> Dict[str, object]:
    """Peer reachability + NVLink bandwidth estimate for ``utils.selftest`` (pod_test analogue)."""
    import time

    from ..ops._build import load

    C = load()
    n = 256 << 20
    buf = SymmetricBuffer(n // 2, torch.bfloat16, env.device, None, env.world)
    buf.tensor.fill_(float(env.rank + 1))
    torch.cuda.synchronize()
    dist.barrier()
    out = {"peers_mapped": len(buf.ptrs), "multicast": bool(getattr(buf, "multicast_ptr", 0))}
    dst = torch.empty(n // 2, dtype=torch.bfloat16, device=env.device)
    peer = (env.rank + 1) % env.world
    C.p2p_copy(dst.data_ptr(), buf.ptrs[peer], n)
    torch.cuda.synchronize()
    out["peer_read_ok"] = bool(torch.all(dst == float(peer + 1)).item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        C.p2p_copy(dst.data_ptr(), buf.ptrs[peer], n)
    e1.record()
    e1.synchronize()
    out["peer_read_gbs"] = round(5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    dist.barrier()
    return out
