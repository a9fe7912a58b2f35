"""Training driver -- the ``main_zero.py`` of this framework.

CLI surface of the reference (``main_zero.py:41-55``): ``--cfg``, ``--model-cfg``,
``--resume``; plus ``key=value`` overrides and ``--synthetic``.  The loop mirrors
the reference hot loop (``main_zero.py:462-588``): per step reshape the host batch
into ``train_context`` rows and ``gradient_accumulation_steps`` micro-batches,
forward/backward each, one ZeRO-1 optimizer update, metrics; every
``evaluation_frequency`` steps run up to ``maximum_evaluation_steps`` validation
batches and write a checkpoint (params + optimizer, ``keep=5``).

``Trainer`` is also the public API that ``bench.py`` drives end to end.
"""
from __future__ import annotations

import argparse
import math
import os
from typing import Dict, Iterable, Iterator, Optional

import numpy as np
import torch

from .config import DEFAULT_CFG, DEFAULT_MODEL_CFG, Config, flatten_dict, load_config
from .data.pipeline import (DevicePrefetcher, StreamingTokenDataset, SyntheticTokenDataset, batch_iterator, rechunk,
                            skip_batches, to_micro_batches)
from .models.gpt import model_getter
from .optim.schedule import schedule_from_config
from .parallel import dist as pdist
from .parallel.zero1 import Zero1Engine
from .utils import checkpoint as ckpt
from .utils.metrics import MetricsSink, StepTimer, compute_tokens_seen, get_logger, mfu


def parse(argv=None):
    p = argparse.ArgumentParser(description="ZeRO-1 GPT pretraining on B200")
    p.add_argument("--cfg", default=DEFAULT_CFG)
    p.add_argument("--model-cfg", default=DEFAULT_MODEL_CFG)
    p.add_argument("--resume", default=False, action="store_true")
    p.add_argument("--synthetic", default=False, action="store_true", help="random-token data source")
    p.add_argument("--max-steps", type=int, default=None, help="stop after this many optimizer steps (this invocation)")
    p.add_argument("--device", default=None, help="cpu | cuda (default: cuda when available)")
    p.add_argument("overrides", nargs="*", help="config overrides, e.g. training.batch_size=32 model.size=test")
    return p.parse_args(argv)


class Trainer:
    def __init__(self, cfg: Config, model_cfg_path: str = DEFAULT_MODEL_CFG, env: Optional[pdist.DistEnv] = None,
                 resume: bool = False, synthetic: Optional[bool] = None, log_metrics: bool = True):
        self.cfg = cfg
        self.env = env or pdist.setup_dp()
        env = self.env
        t, d = cfg.training, cfg.data
        self.log = get_logger(is_main=env.is_main)
        self.synthetic = (d.get("corpus", "synthetic") == "synthetic") if synthetic is None else synthetic
        self.accum = int(t.gradient_accumulation_steps)
        self.train_context = int(t.train_context)
        self.max_context = int(d.max_context)
        if int(t.batch_size) % env.world != 0:
            raise ValueError(f"training.batch_size={t.batch_size} must be divisible by world size {env.world}")
        self.rows_per_rank = int(t.batch_size) // env.world           # rows of max_context tokens
        self.ctx_rows_per_rank = self.rows_per_rank * (self.max_context // self.train_context)
        if self.ctx_rows_per_rank % self.accum != 0:
            raise ValueError("per-rank rows do not divide into gradient_accumulation_steps micro-batches")
        self.global_tokens_per_step = int(t.batch_size) * self.max_context

        torch.manual_seed(int(t.get("seed", 0)))  # same init on every rank (params are replicated)
        on_gpu = env.device.type == "cuda"
        # fp32 init (reference builds the model with dtype=float32, main_zero.py:203-205); the engine
        # keeps the fp32 master shard and re-points the model at the bf16 working copy.
        self.model, self.model_cfg = model_getter(cfg.model.size, model_cfg_path, return_cfg=True,
                                                  dtype=torch.float32, device=env.device)
        self.lr_fn = schedule_from_config(t)
        gd = {"float32": torch.float32, "bfloat16": torch.bfloat16}.get(str(t.get("grad_dtype", "")), None)
        self.engine = Zero1Engine(
            self.model, lr_fn=self.lr_fn, weight_decay=float(t.weight_decay),
            betas=(float(t.get("beta1", 0.9)), float(t.get("beta2", 0.95))), eps=float(t.get("adam_eps", 1e-8)),
            clip=float(t.get("grad_clip", 1.0)), grad_accum_steps=self.accum,
            backend=str(t.get("comm_backend", "auto")),
            compute_dtype=torch.bfloat16 if on_gpu else torch.float32,
            grad_dtype=(gd if on_gpu else torch.float32),
            ag_gemm=bool(t.get("ag_gemm", True)),
            comm_sms=t.get("comm_sms", "auto"),
            clip_mode=str(t.get("grad_clip_mode", "elementwise")),
        )
        self.log.debug(self.engine.memory_report())
        self.log.debug("model %s: %.1fM parameters; optimizer state sharded over %d rank(s); params/grads/activations replicated",
                       cfg.model.size, self.model.cfg.num_params() / 1e6, env.world)

        root = d.get("bucket_path", "") or "."
        self.workdir = os.path.join(root, d.checkpoint_directory)
        self.resume_step = 0
        if resume:
            self.resume_step = ckpt.resume(self.engine, self.workdir)
            self.log.debug("resuming from step %d", self.resume_step)
        elif on_gpu:
            from .ops import cuda as K

            K.manual_seed(int(t.get("seed", 0)) * 1000003 + 0x5EED5EED, env.device)  # dropout streams follow training.seed
        elif cfg.model.get("warm_init", False):
            ckpt.warm_start(self.engine, os.path.join(root, cfg.model.warm_init_dir), bool(cfg.model.get("extend_depth", False)))
            self.log.debug("warm-started from %s", cfg.model.warm_init_dir)
        self.sink = None
        if log_metrics:
            run_cfg = flatten_dict(cfg.to_dict())
            run_cfg.update({f"model.{k}": v for k, v in self.model_cfg.items()})
            run_cfg.update({"training.local_batch_size": self.rows_per_rank, "runtime": f"b200x{env.world}",
                            "Total Training Tokens": int(t.total_steps) * self.global_tokens_per_step,
                            "Total Devices": env.world})
            self.sink = MetricsSink(self.workdir, run_cfg, project=str(d.get("wandb_project", "") or ""), enabled=env.is_main,
                                    run_id=ckpt.stable_run_id(self.workdir) if env.is_main else None)
        if on_gpu and int(t.get("ce_chunk_rows", 0) or 0) > 0:
            from .ops import cuda as K

            K.CONFIG.ce_chunk_rows = int(t.ce_chunk_rows)
        self._timer = StepTimer(env.device)
        self._nvtx = on_gpu and os.environ.get("ZTB_NVTX", "0") == "1"   # NVTX ranges for nsys/ncu timelines
        # training.cuda_graph (or ZTB_CUDA_GRAPH=1): capture one whole optimizer step (all micro-batches, the per-bucket
        # updates on the communication stream, the all-gather waits) into a CUDA graph after `graph_warmup` eager steps
        # and replay it from then on -- one launch per step instead of ~1100, no inter-kernel launch gaps.
        env_graph = os.environ.get("ZTB_CUDA_GRAPH")
        self.use_graph = on_gpu and (bool(t.get("cuda_graph", False)) if env_graph is None else env_graph == "1")
        self.graph_warmup = int(t.get("cuda_graph_warmup", 3))
        self._graph = None
        self._eager_steps = 0

    # ------------------------------------------------------------------ data
    def _samples(self, split: str) -> Iterable[np.ndarray]:
        d, env = self.cfg.data, self.env
        shuffle = int(d.get("shuffle_buffer", 100000))
        if split != "train":
            # validation is re-opened at every evaluation: keep its shuffle buffer (and the refill from disk) small
            shuffle = min(shuffle, int(d.get("validation_shuffle_buffer", 2048)))
        if self.synthetic:
            return SyntheticTokenDataset(self.model.cfg.vocab_size, self.max_context, rank=env.rank,
                                         seed=int(self.cfg.training.get("seed", 0)) + (0 if split == "train" else 7919))
        idx = d.index_path_train if split == "train" else d.index_path_validation
        return StreamingTokenDataset(idx, max_context=self.max_context, rank=env.rank, world=env.world,
                                     shuffle_buffer=shuffle,
                                     seed=23 + self.resume_step,  # reference: Random(23 + resume_step)
                                     epochs=int(self.cfg.training.max_epochs) if split == "train" else 1)

    def train_batches(self) -> DevicePrefetcher:
        it = batch_iterator(self._samples("train"), self.rows_per_rank)
        if self.resume_step > 0:
            # fast-forward the stream to where the interrupted run stopped (reference main_zero.py:437,470-471:
            # resume_step % steps_per_epoch batches are skipped); the synthetic source has no epochs
            spe = int(self.cfg.data.get("steps_per_epoch", 0) or 0)
            n = self.resume_step if (self.synthetic or spe <= 0) else self.resume_step % spe
            it = skip_batches(it, n)
        return DevicePrefetcher(it, self.env.device)

    def _agree_more(self, have: bool) -> bool:
        """True while EVERY rank still has a batch.  Real shards give ranks different batch counts (striding +
        drop_last); the first rank to run dry must stop all of them, or the others spin on its GRAD_READY flags."""
        if self.env.world == 1 or self.synthetic:
            return have
        flag = torch.tensor([1 if have else 0], dtype=torch.int32, device=self.env.device)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        return bool(flag.item())

    def val_batches(self) -> Iterator[torch.Tensor]:
        rows = max(self.rows_per_rank // 4, 1)  # reference: validation batch = batch_size // 4
        return DevicePrefetcher(batch_iterator(self._samples("validation"), rows), self.env.device)

    # ------------------------------------------------------------------ steps
    def train_step(self, batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One optimizer update from a ``(rows_per_rank, max_context)`` token
        batch already on the device.  Returns device tensors (no host sync)."""
        if self.use_graph:
            if self._graph is not None:
                return self._replay(batch)
            if self._eager_steps >= self.graph_warmup:
                self._capture(batch)
                return self._replay(batch)
            self._eager_steps += 1
        return self._train_step_eager(batch)

    # -- CUDA graph --------------------------------------------------------
    def _capture(self, batch: torch.Tensor) -> None:
        from .ops import cuda as K

        eng = self.engine
        eng.synchronize()                      # no pending cross-step waits may leak into the capture
        eng.enable_graph_mode()
        self._g_in = batch.clone()
        torch.cuda.synchronize(self.env.device)
        before = K.LAUNCHES.count
        snap = (eng.step_count, eng._fused.epoch)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            out = self._train_step_eager(self._g_in, in_capture=True)
        self._g_loss = out["train/loss"]
        self._g_launches = K.LAUNCHES.count - before
        K.LAUNCHES.count = before              # the capture pass launched nothing
        # the captured body ran the host-side bookkeeping once without executing anything: rewind it
        eng.step_count, eng._fused.epoch = snap
        self.log.debug("captured one optimizer step as a CUDA graph (%d kernel launches)", self._g_launches)

    def _replay(self, batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        from .ops import cuda as K

        eng = self.engine
        self._g_in.copy_(batch, non_blocking=True)
        lr = eng.stage_step_scalars()
        self._graph.replay()
        eng.after_graph_replay()
        K.LAUNCHES.count += self._g_launches
        return {"train/loss": self._g_loss, "lr": lr}

    def _train_step_eager(self, batch: torch.Tensor, in_capture: bool = False) -> Dict[str, torch.Tensor]:
        self.model.train()
        nvtx = self._nvtx
        if nvtx:
            torch.cuda.nvtx.range_push("train_step")
        mbs = to_micro_batches(rechunk(batch, self.train_context), self.accum)
        self.engine.zero_grad(lazy=True)
        loss_sum = None
        for i in range(self.accum):
            loss = self.model.loss(mbs[i], train=True)
            self.engine.backward(loss, last=(i == self.accum - 1))
            loss_sum = loss.detach() if loss_sum is None else loss_sum + loss.detach()
        if nvtx:
            torch.cuda.nvtx.range_push("optimizer_step")
        lr = self.engine.step()
        if in_capture:
            # self-contained graph: join the communication stream and consume every bucket's "ready" signal here instead
            # of in the next step's forward hooks (costs the overlap of the last buckets' update with the next forward)
            self.engine._fused.wait_all()
            torch.cuda.current_stream(self.env.device).wait_stream(self.engine._fused.stream)
        if nvtx:
            torch.cuda.nvtx.range_pop()
            torch.cuda.nvtx.range_pop()
        mean_loss = loss_sum / self.accum
        return {"train/loss": mean_loss, "lr": lr}

    @torch.no_grad()
    def eval_step(self, batch: torch.Tensor) -> torch.Tensor:
        self.model.eval()
        rows = rechunk(batch, self.train_context)
        return self.model.loss(rows, train=False).detach()

    def evaluate(self) -> Dict[str, float]:
        t = self.cfg.training
        losses = []
        batches = self.val_batches()
        try:
            for _ in range(int(t.maximum_evaluation_steps)):
                b = next(batches, None)
                if not self._agree_more(b is not None):
                    break
                losses.append(self.eval_step(b).float())
        finally:
            batches.close()   # stops the producer thread and releases the dataset generator + its shuffle buffer
        if not losses:
            return {}
        loss = pdist.all_reduce_mean(torch.stack(losses).mean().reshape(1).to(self.env.device)).item()
        return {"validation/loss": loss, "validation/ppl": math.exp(min(loss, 30.0))}

    # ------------------------------------------------------------------ loop
    def fit(self, max_steps: Optional[int] = None) -> Dict[str, float]:
        t = self.cfg.training
        env = self.env
        total_steps = int(t.total_steps)
        eval_every = int(t.evaluation_frequency)
        last: Dict[str, float] = {}
        step = self.resume_step
        new_steps = 0
        flops_tok = self.model.cfg.flops_per_token(self.train_context)
        batches = self.train_batches()
        while True:
            if step >= total_steps or (max_steps is not None and new_steps >= max_steps):
                break
            batch = next(batches, None)
            if not self._agree_more(batch is not None):
                break
            self._timer.start()
            out = self.train_step(batch)
            ms = self._timer.stop_ms()
            step += 1
            new_steps += 1
            loss = pdist.all_reduce_mean(out["train/loss"].float().reshape(1).clone()).item()  # D2H sync, as in the reference (:509-512)
            tok_s = self.global_tokens_per_step / (ms * 1e-3)
            last = {
                "train/loss": loss, "train/ppl": math.exp(min(loss, 30.0)),
                "Train Sequence Length": self.train_context, "Learning Rate": out["lr"],
                "Tokens Seen (B)": compute_tokens_seen(step, int(t.batch_size), self.max_context) / 1e9,
                "perf/step_ms": ms, "perf/tokens_per_s": tok_s,
                "perf/mfu": mfu(tok_s / env.world, flops_tok) if env.device.type == "cuda" else 0.0,
            }
            do_eval = eval_every > 0 and ((step % eval_every == 0) or (step == 1 and bool(t.get("eval_at_step0", False))))
            if do_eval:
                last.update(self.evaluate())
                ckpt.save_checkpoint(self.engine, self.workdir, step, keep=int(self.cfg.data.get("keep_checkpoints", 5)))
            if self.sink:
                self.sink.log(last, step)
            if env.is_main and (step % 10 == 0 or new_steps <= 3 or do_eval):
                self.log.debug("step %d loss %.4f lr %.3e %.1f ms %.0f tok/s%s", step, loss, out["lr"], ms, tok_s,
                               f" val {last['validation/loss']:.4f}" if "validation/loss" in last else "")
        batches.close()
        self.final_step = step
        return last

    def close(self):
        if self.sink:
            self.sink.close()


def main(argv=None):
    args = parse(argv)
    cfg = load_config(args.cfg, args.overrides)
    env = pdist.setup_dp(device=args.device)
    if not args.resume and env.is_main:
        root = cfg.data.get("bucket_path", "") or "."
        ckpt.wipe_checkpoints(os.path.join(root, cfg.data.checkpoint_directory))  # reference :326-342
    pdist.barrier()
    tr = Trainer(cfg, args.model_cfg, env, resume=args.resume, synthetic=True if args.synthetic else None)
    try:
        last = tr.fit(args.max_steps)
        if env.is_main:
            tr.log.debug("done at step %d: %s", tr.final_step, {k: round(v, 4) if isinstance(v, float) else v for k, v in last.items()})
    finally:
        tr.close()
        pdist.shutdown()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
