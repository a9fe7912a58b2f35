"""ZeRO stage-1 engine: replicated low-precision parameters for compute,
fp32 master weights + Adam moments sharded 1/world per rank.

What the reference does (``main_zero.py:443-460,495-500``,
``src/partitioning/xmap_train_functions.py:26-123``): ``pmean`` the full gradient
(all-reduce), slice it to per-device shards, run AdamW on 1/n of every tensor,
and let XLA all-gather the updated *fp32* parameters.  What this engine does per
optimizer step:

    reduce-scatter(grads)  ->  clip/scale + AdamW on the owned shard (fp32
    master, m, v)  ->  write the updated shard in the compute dtype  ->
    all-gather(params)

with three interchangeable communication backends:

``fused``      hand-written sm_100a kernels over NVLink peer memory: the
               reduce-scatter is fused with cast/scale/clip/AdamW and emits the
               bf16 shard; the all-gather is a P2P pull (optionally fused into
               the first consuming GEMM).  No NCCL call.  (``parallel/fused.py``)
``nccl``       ``reduce_scatter_tensor`` + ``all_gather_into_tensor`` per region
               (works on gloo for the CPU plumbing tests as well).
``allreduce``  the reference's own pattern: all-reduce, slice, all-gather.

The engine owns the flat buffers; model parameters (and their ``.grad``) are
re-pointed to views of them, so backward accumulates straight into the
communication buffer (no copy-in), and ``step()`` never walks Python lists of
tensors on the hot path.
"""
from __future__ import annotations
import os
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist

from ..ops import reference
from . import partition


def _dist_info(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class Zero1Engine:
    def __init__(
        self,
        model: torch.nn.Module,
        *,
        lr_fn: Callable[[int], float],
        weight_decay: float = 0.1,
        betas=(0.9, 0.95),
        eps: float = 1e-8,
        clip: Optional[float] = 1.0,
        grad_accum_steps: int = 1,
        group=None,
        backend: str = "auto",
        compute_dtype: Optional[torch.dtype] = None,
        grad_dtype: Optional[torch.dtype] = None,
        decay_mask: Optional[Dict[str, bool]] = None,
        bucket_of=None,   # None: the rule table of parallel/partition.py
        ag_gemm: bool = True,
        comm_sms="auto",   # accepted for config compatibility; SM reservation was measured and dropped (round 1)
        clip_mode: str = "elementwise",
    ):
        self.model = model
        self.group = group
        self.ag_gemm = bool(ag_gemm)
        self.rank, self.world = _dist_info(group)
        # How the per-bucket reduce-scatter+AdamW kernels share the GPU with the backward pass that is still running
        # when they start.  A tcgen05 GEMM CTA owns its SM's shared memory and most of its registers, so a *wide* update
        # kernel can only alternate with the GEMMs (measured round 1: 7.4 ms per step exposed on 1 GPU, and reserving
        # whole SMs for it lost more than that).  "background" launches one 128-thread CTA per SM instead; the GEMM and
        # attention kernels cap their registers so that this CTA is co-resident, and the update hides under backward.
        # Buckets that become ready with less than `bg_min_blocks_left` blocks of backward still to run use the wide form.
        self.update_mode = os.environ.get("ZTB_UPDATE_MODE", "background")
        self.bg_min_blocks_left = int(os.environ.get("ZTB_BG_MIN_BLOCKS", "2"))
        self.lr_fn = lr_fn
        self.weight_decay = float(weight_decay)
        self.beta1, self.beta2 = float(betas[0]), float(betas[1])
        self.eps = float(eps)
        self.clip = clip
        # "elementwise": clamp every gradient element to [-clip, clip] (optax.clip, reference main_zero.py:161),
        # single pass, fused into the update kernel.  "global_norm": scale by min(1, clip / ||g||_2) -- needs the
        # norm of the fully reduced gradient first, so the fused backend runs reduce-scatter and AdamW as two phases.
        if clip_mode not in ("elementwise", "global_norm"):
            raise ValueError(f"unknown clip_mode '{clip_mode}'")
        self.clip_mode = clip_mode
        self.last_grad_norm: Optional[torch.Tensor] = None
        self.grad_accum_steps = int(grad_accum_steps)
        self.step_count = 0  # optimizer updates applied so far

        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("model has no trainable parameters")
        dev = named[0][1].device
        self.device = dev
        src_dtype = named[0][1].dtype
        self.compute_dtype = compute_dtype or (torch.bfloat16 if dev.type == "cuda" else src_dtype)
        self.grad_dtype = grad_dtype or self.compute_dtype
        # grad_dtype != compute_dtype = fp32 "main gradients" next to bf16 parameters.  autograd insists that
        # param.grad has the parameter's dtype, so the fp32 flat buffer is NOT exposed as .grad: every parameter gets a
        # ``_ztb_main_grad`` view instead, which the op layer's fused gradient writers target (wgrad GEMM epilogue with an
        # fp32 C, LayerNorm / embedding backward).  The reference accumulates in bf16 (xmap_train_functions.py:62-70);
        # this is the SURVEY 7.1 option for long accumulation chains.
        self.main_grads = self.grad_dtype != self.compute_dtype
        if self.main_grads and (self.grad_dtype != torch.float32 or dev.type != "cuda"):
            raise ValueError("main gradients are fp32 next to bf16 parameters on a CUDA device")
        if decay_mask is None and hasattr(model, "decay_mask"):
            decay_mask = model.decay_mask()
        # ordered regex rules -> (spec, bucket, weight decay) per parameter: the reference's set_partitions_zero / create_opt_spec
        self.layout = partition.build_layout([(n, tuple(p.shape)) for n, p in named], self.world, decay_mask, bucket_of)
        L = self.layout

        if backend == "auto":
            backend = "fused" if dev.type == "cuda" else "nccl"
        if backend not in ("fused", "nccl", "allreduce"):
            raise ValueError(f"unknown comm backend '{backend}'")
        self.backend = backend
        self._fused = None
        if backend == "fused":
            from .fused import FusedComm  # needs the CUDA extension; fails loudly without it

            self._fused = FusedComm(L, self.rank, self.world, group, dev, self.compute_dtype, self.grad_dtype)
            self.flat_param = self._fused.flat_param
            self.flat_grad = self._fused.flat_grad
        else:
            self.flat_param = torch.zeros(L.total, dtype=self.compute_dtype, device=dev)
            self.flat_grad = torch.zeros(L.total, dtype=self.grad_dtype, device=dev)

        # fp32 master shard taken from the (ideally fp32-initialised) model, then the
        # model is re-pointed at the low-precision flat buffer.
        self.master = torch.zeros(L.shard_total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.decay_vec = L.decay_mask_shard(dev)
        with torch.no_grad():
            for n, p in named:
                r = L.by_name[n]
                full = torch.zeros(r.padded, dtype=torch.float32, device=dev)
                full[: r.numel].copy_(p.detach().reshape(-1).float())
                s = self.rank * r.chunk
                self.master[r.shard_offset: r.shard_offset + r.chunk].copy_(full[s: s + r.chunk])
                self.flat_param[r.offset: r.offset + r.padded].copy_(full.to(self.compute_dtype))
                p.data = L.view(self.flat_param, n)
                if self.main_grads:
                    p.grad = None
                    p._ztb_main_grad = L.view(self.flat_grad, n)
                else:
                    p.grad = L.view(self.flat_grad, n)
        self._named = named
        self._dyn = None
        # Lazy zero_grad (first micro-batch overwrites instead of accumulating) is only sound when every parameter's
        # gradient is produced by a fused writer that honours the overwrite mode.  The op layer falls back to plain
        # autograd accumulation for shapes its kernels do not take (last dim or vocabulary not a multiple of 8); with such
        # a parameter the engine keeps the memset.
        self._lazy_ok = all(p.shape[-1] % 8 == 0 and (p.ndim < 2 or p.shape[0] % 8 == 0) for _, p in named)
        if self.main_grads and not self._lazy_ok:
            raise ValueError("fp32 main gradients need every parameter to have a fused gradient writer (dims multiple of 8)")
        self._grad_shard = torch.zeros(L.shard_total, dtype=self.grad_dtype, device=dev) if backend != "fused" else None
        # fused backend: overlap machinery (see parallel/fused.py)
        self._overwrite_next = False   # first micro-batch of the step overwrites the gradient buffers
        self._last_micro = True
        self._updated = set()
        self._embed_bucket = next((b.index for b in L.buckets if b.name == "embed"), None)
        self._block_bucket = {}
        for b in L.buckets:
            if b.name.startswith("block"):
                self._block_bucket[int(b.name[5:])] = b.index
        self._ag_fused_buckets = ()
        if self._fused is not None and hasattr(model, "block_backward_done_hook"):
            model.block_backward_done_hook = self._on_block_backward_done
            model.block_forward_pre_hook = self._on_block_forward_pre
            self._install_ag_gemm(model)

    # ------------------------------------------------------------------
    @property
    def lr(self) -> float:
        return float(self.lr_fn(self.step_count))

    # -- fused-backend hooks ---------------------------------------------
    def _adam_kwargs(self):
        kw = dict(master=self.master, m=self.exp_avg, v=self.exp_avg_sq, lr=self.lr, beta1=self.beta1, beta2=self.beta2,
                  eps=self.eps, wd=self.weight_decay, step=self.step_count + 1, clip=self.clip or 0.0,
                  grad_scale=1.0 / (self.world * self.grad_accum_steps))
        if self._dyn is not None:
            kw["dyn"] = self._dyn     # CUDA-graph mode: lr and the bias corrections are read from the device
        return kw

    # -- CUDA-graph support ------------------------------------------------
    def enable_graph_mode(self) -> None:
        """Make one optimizer step replayable as a CUDA graph: everything that changes from step to step and used to be
        a launch argument (learning rate, Adam bias corrections, the flag protocol's step number) is read from device
        memory instead.  Call ``stage_step_scalars()`` before each replay and ``after_graph_replay()`` after it."""
        if self._fused is None:
            raise RuntimeError("graph mode needs the fused backend")
        if self._two_phase:
            raise RuntimeError("graph mode does not support clip_mode=global_norm (host-visible all-reduce inside the step)")
        self._dyn = torch.zeros(3, dtype=torch.float32, device=self.device)
        # ring of pinned staging slots: the host may run several replays ahead of the device, and an async copy reads its
        # pinned source when it EXECUTES -- a slot is rewritten only after the copy that last used it has run
        self._dyn_host = [torch.zeros(3, dtype=torch.float32).pin_memory() for _ in range(8)]
        self._dyn_ev = [None] * 8
        self._dyn_i = 0
        self._fused.set_graph_mode(True)
        self.stage_step_scalars()

    def stage_step_scalars(self) -> float:
        """Upload (lr, 1/(1-b1^t), 1/(1-b2^t)) of the update about to run; stream-ordered, no host sync."""
        t = self.step_count + 1
        lr = self.lr
        i = self._dyn_i
        self._dyn_i = (i + 1) % len(self._dyn_host)
        if self._dyn_ev[i] is not None:
            self._dyn_ev[i].synchronize()
        h = self._dyn_host[i]
        h[0] = lr
        h[1] = 1.0 / (1.0 - self.beta1 ** t)
        h[2] = 1.0 / (1.0 - self.beta2 ** t)
        self._dyn.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._dyn_ev[i] = ev
        return lr

    def after_graph_replay(self) -> None:
        """Host-side bookkeeping that the captured ``step()`` did once at capture time."""
        self.step_count += 1
        self._fused.epoch += 1

    def _on_block_backward_done(self, i: int) -> None:
        """Block ``i``'s backward just finished on the last micro-batch: its bucket's gradients are
        final, so start reduce-scatter+AdamW for it now, overlapping the rest of backward."""
        if not self._last_micro or not self._in_engine_backward:
            return
        b = self._block_bucket.get(i)
        if b is not None and b not in self._updated:
            if self._two_phase:
                self._fused.bucket_reduce(b)
            else:
                bg = self.update_mode == "background" and i >= self.bg_min_blocks_left
                self._fused.bucket_update(b, background=bg, **self._adam_kwargs())
            self._updated.add(b)

    def _on_block_forward_pre(self, i: int) -> None:
        b = self._block_bucket.get(i)
        if b is not None and not self._fused._ag_pending[b]:
            self._fused.wait_bucket(b)

    def _install_ag_gemm(self, model) -> None:
        """Tag every block's QKV weight -- the first GEMM that consumes the block's bucket -- so that
        ``ops.linear`` runs the all-gather-fused GEMM when that bucket still has to be gathered."""
        if self.world == 1 or not self.ag_gemm or os.environ.get("ZTB_AG_GEMM", "1") != "1":
            return
        fused = []
        for i, blk in enumerate(getattr(model, "blocks", [])):
            b = self._block_bucket.get(i)
            name = f"blocks.{i}.attn.qkv.weight"
            if b is None or name not in self.layout.by_name:
                continue
            if not self._fused.ag_plan(b, name)["ok"]:
                continue
            blk.attn.qkv.weight._ztb_ag = _AgHandle(self._fused, b, name)
            fused.append(b)
        self._ag_fused_buckets = tuple(fused)

    _in_engine_backward = False

    def backward(self, loss: torch.Tensor, last: bool = True) -> None:
        """``loss.backward()`` with the engine's gradient-buffer protocol: the first micro-batch after
        ``zero_grad()`` overwrites the flat gradient buffer (no memset), later ones accumulate, and on
        the ``last`` micro-batch every bucket's optimizer update starts as soon as its block's backward
        is done.  Plain ``loss.backward()`` + ``step()`` remains valid (no overlap)."""
        if self._fused is None:
            loss.backward()
            self._fold_stray_grads()
            return
        from ..ops import cuda as K

        self._last_micro = bool(last)
        if last and not self._updated:
            self._fused.begin_step()
        self._in_engine_backward = True
        K.GRAD_MODE["accumulate"] = not self._overwrite_next
        try:
            loss.backward()
        finally:
            K.GRAD_MODE["accumulate"] = True
            K.GRAD_MODE["lazy_zero"] = False
            self._in_engine_backward = False
        self._overwrite_next = False
        self._fold_stray_grads()

    def _fold_stray_grads(self) -> None:
        """Main-gradient mode: a gradient that reached a parameter through plain autograd (an op without a fused
        writer) sits in ``.grad`` in the compute dtype -- add it to the fp32 buffer and drop it."""
        if not self.main_grads:
            return
        for _, p in self._named:
            if p.grad is not None:
                p._ztb_main_grad.add_(p.grad.to(p._ztb_main_grad.dtype))
                p.grad = None

    def zero_grad(self, lazy: bool = False) -> None:
        """``lazy=True`` (fused backend, gradients produced through ``engine.backward``): skip the
        memset, the next backward overwrites instead of accumulating."""
        if lazy and self._fused is not None and self._lazy_ok:
            self._overwrite_next = True
            from ..ops import cuda as K

            K.GRAD_MODE["lazy_zero"] = True    # seen by gradient writers that run before engine.backward (chunked LM head)
            return
        self.flat_grad.zero_()
        if self.main_grads:
            return
        # autograd may have replaced .grad if someone set it to None; re-point cheaply
        for n, p in self._named:
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + self.layout.by_name[n].offset * self.flat_grad.element_size():
                p.grad = self.layout.view(self.flat_grad, n)

    # ------------------------------------------------------------------
    def step(self) -> float:
        """Apply one optimizer update from the gradients accumulated in
        ``flat_grad`` (sum over micro-batches of per-micro-batch mean-loss
        gradients).  Returns the learning rate used."""
        lr = self.lr
        t = self.step_count + 1
        scale = 1.0 / (self.world * self.grad_accum_steps)
        if self.backend == "fused":
            f = self._fused
            if not self._updated:
                f.begin_step()          # no bucket was started from backward (plain loss.backward())
            kw = self._adam_kwargs()
            if self._two_phase:
                for b in range(f.nb):
                    if b not in self._updated:
                        f.bucket_reduce(b)
                coef = self._global_norm_coef(scale, stream=f.stream)
                kw.update(clip=0.0, reduced=True, scale_mul=coef)
                for b in range(f.nb):
                    f.bucket_update(b, **kw)
            else:
                for b in range(f.nb):       # whatever backward did not start: embed bucket, or everything
                    if b not in self._updated:
                        f.bucket_update(b, **kw)
            self._updated.clear()
            # all-gather: embed bucket (+ anything without a fused consumer) by the standalone P2P pull,
            # queued now so it overlaps the tail of this step; block buckets inside their QKV GEMM
            f.prefetch_all(fused_buckets=self._ag_fused_buckets if f.world > 1 else ())
            if self._embed_bucket is not None:
                f.wait_bucket(self._embed_bucket)
            if not hasattr(self.model, "block_forward_pre_hook") or self.model.block_forward_pre_hook is None:
                f.wait_all()
        else:
            self._reduce_to_shard()
            clip = self.clip
            if self._two_phase:
                scale = scale * float(self._global_norm_coef(scale, shard=self._grad_shard))
                clip = None
            with torch.no_grad():
                reference.adamw_step_(
                    self.master, self._grad_shard, self.exp_avg, self.exp_avg_sq,
                    lr=lr, beta1=self.beta1, beta2=self.beta2, eps=self.eps,
                    wd=self.decay_vec * self.weight_decay, step=t, clip=clip, grad_scale=scale,
                )
                self.layout.scatter_shard_to_flat(self.master.to(self.compute_dtype), self.flat_param, self.rank)
            self._all_gather_params()
        self.step_count += 1
        return lr

    @property
    def _two_phase(self) -> bool:
        return self.clip_mode == "global_norm" and bool(self.clip)

    def _global_norm_coef(self, scale: float, shard: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        """min(1, clip / ||scale * g||_2) as a 1-element fp32 device tensor; ``g`` is the reduced gradient whose
        shards live on the ranks (``shard``, or this rank's chunks of the flat buffer after ``bucket_reduce``)."""
        import contextlib

        ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
        with ctx, torch.no_grad():
            if shard is None:
                shard = self.layout.gather_shard_from_flat(self.flat_grad, self.rank)
            sq = shard.float().pow(2).sum().reshape(1)
            if self.world > 1:
                dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=self.group)
            norm = sq.sqrt() * scale
            self.last_grad_norm = norm
            return torch.clamp(float(self.clip) / (norm + 1e-6), max=1.0)

    # -- unfused collectives --------------------------------------------
    def _reduce_to_shard(self) -> None:
        L = self.layout
        if self.world == 1:
            L.gather_shard_from_flat(self.flat_grad, 0, self._grad_shard)
            return
        if self.backend == "allreduce":
            # reference pattern (xmap_train_functions.py:84 + main_zero.py:497): all-reduce, then local slice
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            L.gather_shard_from_flat(self.flat_grad, self.rank, self._grad_shard)
            return
        for r in L.regions:
            out = self._grad_shard[r.shard_offset: r.shard_offset + r.chunk]
            dist.reduce_scatter_tensor(out, self.flat_grad[r.offset: r.offset + r.padded], op=dist.ReduceOp.SUM, group=self.group)

    def _all_gather_params(self) -> None:
        if self.world == 1:
            return
        L = self.layout
        inplace_ok = self.flat_param.is_cuda  # NCCL supports in-place all-gather; gloo gets a staging copy
        for r in L.regions:
            s, e = r.chunk_range(self.rank)
            src = self.flat_param[s:e] if inplace_ok else self.flat_param[s:e].clone()
            dist.all_gather_into_tensor(self.flat_param[r.offset: r.offset + r.padded], src, group=self.group)

    def synchronize(self) -> None:
        """Make the optimizer state and the parameters safe to read from the current stream / host
        (the fused backend updates them on its own communication stream)."""
        if self._fused is not None:
            self._fused.wait_all()
            torch.cuda.current_stream(self.device).synchronize()

    # -- state -----------------------------------------------------------
    def shard_state(self) -> Dict[str, object]:
        """This rank's optimizer shard (what the reference keeps sharded on
        device, ``main_zero.py:447-450``)."""
        self.synchronize()
        return {
            "step": self.step_count, "rank": self.rank, "world": self.world,
            "master": self.master.detach().cpu(), "exp_avg": self.exp_avg.detach().cpu(),
            "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
        }

    def load_shard_state(self, st: Dict[str, object]) -> None:
        if int(st["world"]) != self.world or int(st["rank"]) != self.rank:
            raise ValueError("shard state belongs to a different (rank, world); load a consolidated state instead")
        self.step_count = int(st["step"])
        self.master.copy_(st["master"])
        self.exp_avg.copy_(st["exp_avg"])
        self.exp_avg_sq.copy_(st["exp_avg_sq"])
        self._push_master_to_params()

    def consolidated_state(self, dst: int = 0) -> Optional[Dict[str, object]]:
        """Gather the sharded state into full per-parameter fp32 tensors on
        ``dst`` (the reference gathers with ``process_allgather`` before saving,
        ``main_zero.py:554-557``).  World-size independent, so a run can resume
        on a different number of GPUs."""
        L = self.layout
        self.synchronize()
        out = {"step": self.step_count, "params": {}, "exp_avg": {}, "exp_avg_sq": {}}
        for key, shard in (("params", self.master), ("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
            if self.world > 1:
                parts = [torch.empty_like(shard) for _ in range(self.world)]
                dist.all_gather(parts, shard, group=self.group)
            else:
                parts = [shard]
            if self.rank == dst or self.world == 1:
                for r in L.regions:
                    full = torch.cat([pt[r.shard_offset: r.shard_offset + r.chunk] for pt in parts])
                    out[key][r.name] = full[: r.numel].view(r.shape).cpu().clone()
        return out if (self.rank == dst or self.world == 1) else None

    def load_consolidated_state(self, st: Dict[str, object], load_moments: bool = True, adopt_step: bool = True) -> None:
        L = self.layout
        with torch.no_grad():
            for key, shard in (("params", self.master), ("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                if key != "params" and (not load_moments or key not in st):
                    continue
                for r in L.regions:
                    full = torch.zeros(r.padded, dtype=torch.float32)
                    full[: r.numel].copy_(st[key][r.name].reshape(-1).float())
                    s = self.rank * r.chunk
                    shard[r.shard_offset: r.shard_offset + r.chunk].copy_(full[s: s + r.chunk])
        if adopt_step:
            self.step_count = int(st.get("step", 0))
        self._push_master_to_params(full_params=st["params"])

    def _push_master_to_params(self, full_params: Optional[Dict[str, torch.Tensor]] = None) -> None:
        with torch.no_grad():
            if full_params is not None:
                for r in self.layout.regions:
                    self.flat_param[r.offset: r.offset + r.numel].copy_(full_params[r.name].reshape(-1).to(self.compute_dtype))
                return
            self.layout.scatter_shard_to_flat(self.master.to(self.compute_dtype), self.flat_param, self.rank)
            if self.backend == "fused":
                self._fused.all_gather_now()
            else:
                self._all_gather_params()

    def full_params_fp32(self, dst: int = 0):
        st = self.consolidated_state(dst)
        return None if st is None else st["params"]

    def memory_report(self) -> str:
        gb = 1.0 / (1 << 30)
        L = self.layout
        pb = L.total * self.flat_param.element_size()
        gbs = L.total * self.flat_grad.element_size()
        ob = 3 * L.shard_total * 4
        return (f"ZeRO-1 rank {self.rank}/{self.world} [{self.backend}]: params {pb * gb:.2f} GiB ({self.compute_dtype}), "
                f"grads {gbs * gb:.2f} GiB ({self.grad_dtype}), fp32 master+m+v shard {ob * gb:.2f} GiB "
                f"(unsharded would be {ob * self.world * gb:.2f} GiB)")


class _AgHandle:
    """Attached to a block's QKV weight: lets the op layer run the all-gather-fused GEMM."""

    def __init__(self, fused, bucket: int, region: str):
        self.fused, self.bucket, self.region = fused, bucket, region

    def pending(self) -> bool:
        return self.fused._ag_pending[self.bucket]

    def gemm(self, x2, w):
        return self.fused.gemm_ag(self.bucket, self.region, x2, w)

    def fallback(self) -> None:
        self.fused.wait_bucket(self.bucket)


class PlainAdamW:
    """Unsharded fp32 AdamW with the same math (elementwise clip -> AdamW with
    ``ndim >= 2`` decay mask).  The oracle for the ZeRO-1 equivalence tests and
    the optimizer of the single-process reference arm."""

    def __init__(self, model, *, lr_fn, weight_decay=0.1, betas=(0.9, 0.95), eps=1e-8, clip=1.0, grad_accum_steps=1, world=1,
                 clip_mode="elementwise"):
        self.clip_mode = clip_mode
        self.params = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.lr_fn = lr_fn
        self.wd, self.b1, self.b2, self.eps, self.clip = weight_decay, betas[0], betas[1], eps, clip
        self.scale = 1.0 / (grad_accum_steps * world)
        self.m = {n: torch.zeros_like(p, dtype=torch.float32) for n, p in self.params}
        self.v = {n: torch.zeros_like(p, dtype=torch.float32) for n, p in self.params}
        self.master = {n: p.detach().float().clone() for n, p in self.params}
        self.step_count = 0

    def zero_grad(self):
        for _, p in self.params:
            p.grad = None

    def step(self) -> float:
        lr = float(self.lr_fn(self.step_count))
        t = self.step_count + 1
        with torch.no_grad():
            scale, clip = self.scale, self.clip
            if self.clip_mode == "global_norm" and clip:
                norm = torch.sqrt(sum((p.grad.float() * scale).pow(2).sum() for _, p in self.params if p.grad is not None))
                scale, clip = scale * min(1.0, float(clip) / (float(norm) + 1e-6)), None
            for n, p in self.params:
                if p.grad is None:
                    continue
                reference.adamw_step_(self.master[n], p.grad, self.m[n], self.v[n], lr=lr, beta1=self.b1, beta2=self.b2,
                                      eps=self.eps, wd=self.wd if p.ndim >= 2 else 0.0, step=t, clip=clip, grad_scale=scale)
                p.copy_(self.master[n].to(p.dtype))
        self.step_count += 1
        return lr
