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

import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

ADAM_TILE = 1024  # elements per CTA tile; must match kAdamTile in csrc/optim.cu
COPY_TILE = 8192  # must match kCopyTile in csrc/comm.cu (and AG_TILE of the 1-CTA all-gather GEMM)
AG2_TILE = 4096   # must match AG_TILE_ELEMS in csrc/gemm2_tcgen05.cu (bulk-copy gather of the CTA-pair all-gather GEMM)
RS_TILE = 2048    # elements per CTA tile of csrc/comm.cu reduce_scatter_kernel
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
    """Per-bucket fused reduce-scatter+AdamW and all-gather over NVLink peer memory.

    Streams: everything here runs on a dedicated communication stream.  ``bucket_update(b)`` is
    called from the backward pass as soon as bucket ``b``'s gradients are final (overlapping the
    remaining backward); ``prefetch_all()`` queues the all-gather pulls in forward order so they
    overlap the next forward; ``wait_bucket(b)`` makes the compute stream wait for bucket ``b``.

    Flags (symmetric int32 pad, one row of ``MAX_RANKS`` per (kind, bucket)) carry the step number:
      GRAD_READY[b][r]   rank r's gradients of bucket b are final      (consumed by every rank's AdamW)
      SHARD_READY[b][r]  rank r has written its updated chunk of b     (consumed by every rank's pull)
    Cross-rank waits run in a one-warp kernel so a slow peer never pins SM resources.
    """

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
        self.nb = len(layout.buckets)
        self._param = SymmetricBuffer(layout.total, torch.bfloat16, device, group, world)
        self._grad = SymmetricBuffer(layout.total, grad_dtype, device, group, world)
        self._flags = SymmetricBuffer(2 * self.nb * MAX_RANKS, torch.int32, device, group, world)
        self.flat_param, self.flat_grad = self._param.tensor, self._grad.tensor
        self.adam, self.pull, self.rs_inplace = [], [], []
        for b in layout.buckets:
            segs = [(r.offset + rank * r.chunk, r.shard_offset, r.chunk, int(r.decay)) for r in b.regions]
            t, tiles = pack_segments(segs, device, ADAM_TILE)
            self.adam.append((t, len(segs), tiles))
            # stand-alone reduce-scatter that leaves the sum in place (own chunk of the local gradient buffer):
            # the two-phase path of global-norm clipping
            rs = [(r.offset + rank * r.chunk, r.offset + rank * r.chunk, r.chunk, 0) for r in b.regions]
            t, tiles = pack_segments(rs, device, RS_TILE)
            self.rs_inplace.append((t, len(rs), tiles))
            # rotated source order: at any moment every rank reads from a different owner, so no
            # owner's NVLink egress is shared by several readers
            pl = [(r.offset + p * r.chunk, 0, r.chunk, p) for p in self._peer_order(rank, world) for r in b.regions]
            t, tiles = pack_segments(pl, device, COPY_TILE)
            self.pull.append((t, len(pl), tiles))
        self.grad_ptrs = self._grad.ptrs[rank:] + self._grad.ptrs[:rank]  # local copy first
        # all-gather fused into the bucket's first GEMM: pull list with that weight's chunks first
        self.ag_lists = {}
        self.ag_done = torch.zeros(8, dtype=torch.int32, device=device)
        self._ag_done_base = 0   # the per-owner counters are cumulative across launches (no memset per GEMM); host mirror of their value
        # CTA-pair (cta_group::2) all-gather GEMM with a bulk-copy (TMA) gather ring; 0 = the 1-CTA kernel whose gather
        # warps move the data through registers (round 1)
        self.ag_two_cta = os.environ.get("ZTB_AG_2CTA", "1") == "1"
        self._ag_pending = [False] * self.nb
        self.epoch = 0
        # CUDA-graph mode: a replayed launch cannot carry a new step number, so the counter also lives on the device and
        # the flag kernels are passed (offset 0, pointer) instead of the value.  ``begin_step`` bumps both copies; the
        # graph replays the device-side bump, the replay wrapper (Zero1Engine.after_graph_replay) the host-side one.
        self.graph_mode = False
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.stream = torch.cuda.Stream(device)
        self.sm_count = torch.cuda.get_device_properties(device).multi_processor_count
        self.ready = [torch.cuda.Event() for _ in range(self.nb)]       # bucket b usable by compute
        self._ev = torch.cuda.Event()
        self._pending = [False] * self.nb
        # NVLS: with a multicast mapping of both flat buffers the update kernel reduces the gradients in the
        # switch (multimem.ld_reduce) and pushes the new bf16 chunk to every rank (multimem.st): reduce-scatter,
        # AdamW and all-gather are one kernel, and the consumer only waits for the SHARD_READY flags.
        self.nvls = bool(world > 1 and os.environ.get("ZTB_NVLS", "1") != "0" and not self.grad_fp32
                         and getattr(self._grad, "multicast_ptr", 0) and getattr(self._param, "multicast_ptr", 0))
        self._pushed = [False] * self.nb      # bucket's new chunk was multicast to every rank by its owner
        self._flagwait = [False] * self.nb    # ... so the consumer waits for the owners' flags instead of pulling
        if world > 1:
            dist.barrier(group)

    @staticmethod
    def _peer_order(rank: int, world: int):
        return [(rank + k) % world for k in range(1, world)]

    # flag addressing ----------------------------------------------------
    def _slot(self, kind: int, bucket: int, src: int) -> int:
        return (kind * self.nb + bucket) * MAX_RANKS + src

    def _my_row(self, kind: int, bucket: int) -> int:
        return self._flags.ptrs[self.rank] + 4 * self._slot(kind, bucket, 0)

    def _epoch_arg(self, value: int):
        """(value, value_ptr) for the flag kernels: the literal step number, or (0, &epoch_dev) in CUDA-graph mode."""
        return (0, self.epoch_dev.data_ptr()) if self.graph_mode else (int(value), 0)

    def _signal(self, kind: int, bucket: int, value: int) -> None:
        v, ptr = self._epoch_arg(value)
        self._C.signal(self._flags.ptrs, self._slot(kind, bucket, self.rank), v, ptr)

    def _wait(self, kind: int, bucket: int, value: int) -> None:
        v, ptr = self._epoch_arg(value)
        self._C.wait_flags(self._my_row(kind, bucket), self.world, v, ptr)

    def set_graph_mode(self, on: bool) -> None:
        if on and not self.graph_mode:
            self.epoch_dev.fill_(self.epoch)
        if bool(on) != self.graph_mode:   # graph mode zeroes the all-gather GEMM's landing counters per launch, eager mode lets them run
            self.ag_done.zero_()
            self._ag_done_base = 0
        self.graph_mode = bool(on)

    # per-step protocol --------------------------------------------------
    def begin_step(self) -> None:
        self.epoch += 1
        if self.graph_mode:
            # on the compute stream: after this step's forward (whose waits still mean the previous step) and before every
            # flag kernel of this step (bucket_update makes the communication stream wait on the compute stream first)
            self.epoch_dev.add_(1)
        self._pushed = [False] * self.nb

    def bucket_update(self, b: int, *, master, m, v, lr, beta1, beta2, eps, wd, step, clip, grad_scale, background: bool = False,
                      reduced: bool = False, scale_mul: Optional[torch.Tensor] = None, dyn: Optional[torch.Tensor] = None) -> None:
        """Bucket ``b``'s gradients are final on the current (compute) stream: reduce-scatter them
        from every peer, apply AdamW to the owned shard and publish the new bf16 chunk.

        ``background``: launch at most one small CTA per SM.  The tcgen05 GEMM / attention kernels cap their
        register use so that such a CTA is co-resident with theirs: the update then runs *under* the rest of the
        backward pass (on its spare memory bandwidth) instead of alternating with it."""
        e = self.epoch
        cur = torch.cuda.current_stream(self.device)
        self._ev.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._ev)
            if self.world > 1 and not reduced:
                self._signal(FLAG_GRAD_READY, b, e)
                self._wait(FLAG_GRAD_READY, b, e)
            segs, n, tiles = self.adam[b]
            # reduced=True: bucket_reduce() already left the cross-rank sum in the local buffer
            ptrs = self._grad.ptrs[self.rank:self.rank + 1] if (reduced or self.world == 1) else self.grad_ptrs
            mc_grad = self._grad.multicast_ptr if (self.nvls and not reduced) else 0
            mc_param = self._param.multicast_ptr if self.nvls else 0
            self._C.adamw_shard(master, m, v, self.flat_param, ptrs, self.grad_fp32, segs, n, tiles, lr, beta1, beta2, eps, wd, clip,
                                grad_scale, step, 0, 0, self.sm_count if background else 0, scale_mul.data_ptr() if scale_mul is not None else 0, mc_grad, mc_param,
                                dyn.data_ptr() if dyn is not None else 0)
            self._pushed[b] = bool(mc_param)
            if self.world > 1:
                self._signal(FLAG_SHARD_READY, b, e)
            else:
                self.ready[b].record(self.stream)
        self._pending[b] = True

    def bucket_reduce(self, b: int) -> None:
        """Phase 1 of a two-phase update (global-norm clipping): reduce-scatter bucket ``b`` over peer memory,
        leaving the sum in this rank's own chunk of the local gradient buffer."""
        if self.grad_fp32:
            raise NotImplementedError("two-phase update (global-norm clipping) needs bf16 gradient buffers")
        e = self.epoch
        self._ev.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._ev)
            if self.world > 1:
                self._signal(FLAG_GRAD_READY, b, e)
                self._wait(FLAG_GRAD_READY, b, e)
                segs, n, tiles = self.rs_inplace[b]
                self._C.reduce_scatter(self.flat_grad, self.grad_ptrs, segs, n, tiles, 0, 0)

    def prefetch_all(self, fused_buckets=()) -> None:
        """Queue the standalone all-gather (forward order) behind the updates for every bucket that
        is not gathered inside its first GEMM (``fused_buckets``)."""
        if self.world == 1:
            return
        e = self.epoch
        for b in range(self.nb):
            if self._pushed[b]:
                self._flagwait[b] = True       # nothing to pull: every owner multicast its chunk
            elif b in fused_buckets:
                self._ag_pending[b] = True
        with torch.cuda.stream(self.stream):
            for b in range(self.nb):
                if b in fused_buckets or self._pushed[b]:
                    continue
                self._wait(FLAG_SHARD_READY, b, e)
                segs, n, tiles = self.pull[b]
                self._C.allgather_pull(self.flat_param, self._param.ptrs, segs, n, tiles, 0, 0)
                self.ready[b].record(self.stream)

    def ag_plan(self, b: int, first_region: str, tile: int = COPY_TILE):
        """Pull list for bucket ``b`` with ``first_region``'s chunks leading (those gate the GEMM)."""
        key = (b, first_region, tile)
        if key not in self.ag_lists:
            regs = self.layout.buckets[b].regions
            lead = [r for r in regs if r.name == first_region]
            rest = [r for r in regs if r.name != first_region]
            assert len(lead) == 1
            r0 = lead[0]
            order = self._peer_order(self.rank, self.world)
            pl = [(r0.offset + p * r0.chunk, 0, r0.chunk, p) for p in order]
            n_b = len(pl)
            pl += [(r.offset + p * r.chunk, 0, r.chunk, p) for p in order for r in rest]
            segs, tiles = pack_segments(pl, self.device, tile)
            k = r0.shape[1]
            ok = r0.padded == r0.numel and (r0.chunk % k == 0)
            self.ag_lists[key] = dict(segs=segs, nseg=len(pl), n_b=n_b, tiles=tiles, b_tiles=(r0.chunk + tile - 1) // tile,
                                      rows_per_rank=r0.chunk // k, ok=ok)
        return self.ag_lists[key]

    def gemm_ag(self, b: int, first_region: str, x2: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        """``x2 @ w.T`` where ``w`` is bucket ``b``'s first-consumed weight: the bucket's all-gather
        runs inside the GEMM kernel (compute stream); marks the bucket as gathered."""
        two_cta = self.ag_two_cta and (plan0 := self.ag_plan(b, first_region, AG2_TILE))["rows_per_rank"] % 128 == 0
        plan = plan0 if two_cta else self.ag_plan(b, first_region)
        out = self._C.gemm_ag(x2, w, self._param.ptrs, self.flat_param, plan["segs"], plan["nseg"], plan["n_b"], plan["tiles"],
                              self._my_row(FLAG_SHARD_READY, b), self._epoch_arg(self.epoch)[0], self.ag_done, plan["b_tiles"], plan["rows_per_rank"],
                              self.world, self.rank, two_cta, self._epoch_arg(self.epoch)[1],
                              -1 if self.graph_mode else self._ag_done_base)
        if not self.graph_mode:   # (a captured launch cannot carry a new base: graph mode zeroes the counters instead)
            self._ag_done_base = (self._ag_done_base + plan["b_tiles"]) & 0xFFFFFFFF
        self._ag_pending[b] = False
        self._pending[b] = False
        return out

    def wait_bucket(self, b: int) -> None:
        """Compute stream waits until bucket ``b`` holds this step's parameters."""
        if self._ag_pending[b]:
            # nobody ran the fused GEMM for this bucket (e.g. state access before the next forward):
            # fall back to the standalone pull on the compute stream
            self._ag_pending[b] = False
            self._wait(FLAG_SHARD_READY, b, self.epoch)
            segs, n, tiles = self.pull[b]
            self._C.allgather_pull(self.flat_param, self._param.ptrs, segs, n, tiles, 0, 0)
            self._pending[b] = False
            return
        if self._flagwait[b]:
            self._flagwait[b] = False
            self._wait(FLAG_SHARD_READY, b, self.epoch)   # one warp, compute stream
            self._pending[b] = False
            return
        if self._pending[b]:
            torch.cuda.current_stream(self.device).wait_event(self.ready[b])
            self._pending[b] = False

    def wait_all(self) -> None:
        for b in range(self.nb):
            self.wait_bucket(b)

    def all_gather_now(self) -> None:
        """Out-of-band all-gather (checkpoint load): publish + pull everything under a fresh epoch."""
        if self.world == 1:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        self.begin_step()
        e = self.epoch
        with torch.cuda.stream(self.stream):
            for b in range(self.nb):
                self._signal(FLAG_GRAD_READY, b, e)
                self._signal(FLAG_SHARD_READY, b, e)
        self._pending = [True] * self.nb
        self.prefetch_all()
        self.wait_all()
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)


def p2p_probe(env) -> Dict[str, object]:
    """Peer reachability + NVLink bandwidth estimate for ``utils.selftest`` (pod_test analogue)."""
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
