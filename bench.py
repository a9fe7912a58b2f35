"""Headline benchmark: tokens/s of ZeRO-1 GPT-2 1.3B pretraining (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One optimizer step = zero grads, ``accum`` micro-batches of forward+backward (dropout 0.1 on,
as in the reference config), reduce-scatter + sharded AdamW + all-gather.  Weak scaling: the
per-GPU batch is fixed.  Two timed regions, both K steps bracketed by barrier + synchronize and
timed with CUDA events (max over ranks):

  value  device arm   -- token batches already resident on the device
  e2e    public API   -- ``Trainer.train_step`` fed from pinned host memory every step (H2D copy
                         inside the timed region) and the loss read back to the host every step

``--impl reference`` must run the unmodified reference: it is JAX/TPU-only, has no
setup.py/pyproject (pip: "not installable") and jax/flax/optax are absent offline, so it
reports ``unavailable``.  ``--impl baseline`` runs the reference-equivalent stock-library build
(``baseline/strong.py``: cuBLAS + flash_attn + fused AdamW + one NCCL reduce-scatter / all-gather) for an
on-box comparison; its committed same-box results (``profiles/baseline_arm.json``) are the denominator
of ``vs_baseline`` (BASELINE.md publishes no throughput for this config, see its section 2).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = os.environ.get("ZTB_BENCH_MODEL", "1_3b")
SEQ = int(os.environ.get("ZTB_BENCH_SEQ", "1024"))
ROWS_PER_GPU = int(os.environ.get("ZTB_BENCH_ROWS", "32"))     # sequences of SEQ tokens per GPU per step
ACCUM = int(os.environ.get("ZTB_BENCH_ACCUM", "2"))            # micro-batches per step
MODEL_LABEL = {"1_3b": "1.3B", "760m": "760M", "417m": "417M", "125m": "125M"}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=8)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference", "baseline"])
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.samples, self._stop, self.t = gpu_index, [], False, None

    def _run(self):
        while not self._stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self) -> dict:
        self._stop = True
        if self.t:
            self.t.join(timeout=6)
        sm = sorted(int(float(s[1])) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit())
        mx = [int(float(s[2])) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def _baseline_number(world: int):
    """Same-box result of the stock-library arm for this (model, shape, GPU count), committed under profiles/."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "baseline_arm.json")))
    except (OSError, ValueError):
        return None
    key = f"{MODEL}/seq{SEQ}/rows{ROWS_PER_GPU}/accum{ACCUM}/n{world}"
    rec = table.get(key)
    if not rec:
        return None
    return {"value": float(rec["value"]), "source": f"profiles/baseline_arm.json[{key}] ({rec.get('note', 'bench.py --impl baseline, same pod')})"}


def main():
    args = parse()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        print(json.dumps({"impl": "reference", "unavailable": "reference is JAX/Flax on TPU only (main_zero.py:200-201 raises on GPU); "
                          "no setup.py/pyproject (pip: not installable) and jax/flax/optax/omegaconf are absent offline"}))
        return 0
    import torch
    import torch.distributed as dist

    from zero_transformer_b200.config import DEFAULT_MODEL_CFG, load_config
    from zero_transformer_b200.parallel import dist as pdist

    env = pdist.setup_dp()
    if env.device.type != "cuda":
        print(json.dumps({"error": "bench.py needs a CUDA device"}))
        return 1
    world = env.world
    assert world == args.gpus or world == 1 and args.gpus == 1, f"launched {world} ranks for --gpus {args.gpus}"
    if args.impl == "baseline":
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        from strong import BaselineTrainer   # stock libraries only: none of this repo's kernels / ops / engine

        tr = BaselineTrainer(MODEL, DEFAULT_MODEL_CFG, env.device, env.rank, world, ACCUM, SEQ)
        ops = None
        model_cfg = None
    else:
        from zero_transformer_b200 import ops
        from zero_transformer_b200.train import Trainer

        over = [
            f"model.size={MODEL}", f"training.batch_size={ROWS_PER_GPU * world}", f"training.train_context={SEQ}",
            f"data.max_context={SEQ}", f"training.gradient_accumulation_steps={ACCUM}", "training.evaluation_frequency=0",
            "data.corpus=synthetic", "data.checkpoint_directory=gpurun_out/bench_ckpt",
            "training.comm_backend=auto", "training.grad_dtype=bfloat16",
        ]
        cfg = load_config(None, over)
        tr = Trainer(cfg, DEFAULT_MODEL_CFG, env, log_metrics=False)
        model_cfg = tr.model.cfg
    if model_cfg is None:   # baseline arm: shape and FLOP count from the same YAML table (no model of ours is built)
        from zero_transformer_b200.config import load_yaml
        from zero_transformer_b200.models.gpt import _cfg_from_entry

        model_cfg = _cfg_from_entry(load_yaml(DEFAULT_MODEL_CFG)[MODEL], None)
    vocab = model_cfg.vocab_size
    tokens_per_step = ROWS_PER_GPU * world * SEQ

    # ---------------- data: pinned host batches (e2e arm) and resident device batches (device arm)
    g = torch.Generator().manual_seed(1234 + env.rank)
    n_host = 4
    host = [torch.randint(0, vocab, (ROWS_PER_GPU, SEQ), dtype=torch.int32, generator=g).pin_memory() for _ in range(n_host)]
    dev = [h.to(env.device) for h in host]
    h2d_bytes = host[0].numel() * host[0].element_size()
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def step_device(i):
        return tr.train_step(dev[i % n_host])

    def step_e2e(i):
        b = host[i % n_host].to(env.device, non_blocking=True)   # H2D from pinned memory, every step
        out = tr.train_step(b)
        loss_host.copy_(out["train/loss"].float().reshape(1), non_blocking=False)  # D2H read of the step's result
        return float(loss_host[0])

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(steps):
            last = fn(i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        return pdist.max_over_ranks(ms, env.device), last

    n_warm = max(args.warmup, 3)
    if getattr(tr, "use_graph", False):
        n_warm = max(n_warm, tr.graph_warmup + 2)   # the eager warm-up steps, the capture itself and one replay stay out of the timed region
    for i in range(n_warm):
        step_e2e(i)
    counter = getattr(ops._k(), "LAUNCHES", None) if args.impl == "ours" else None
    sampler = ClockSampler(env.local_rank)
    if env.is_main:
        sampler.start()
    if counter is not None:
        counter.reset()
    ms_dev, _ = timed(step_device, args.steps)
    launches = counter.count if counter is not None else 0
    ms_e2e, last_loss = timed(step_e2e, args.steps)
    clocks = sampler.stop() if env.is_main else {}

    if env.is_main:
        tok_s = tokens_per_step * args.steps / (ms_dev * 1e-3)
        tok_s_e2e = tokens_per_step * args.steps / (ms_e2e * 1e-3)
        flops_tok = model_cfg.flops_per_token(SEQ)            # causal attention counted once (lower triangle)
        flops_tok_dense = model_cfg.flops_per_token(SEQ, causal=False)
        base = _baseline_number(world)
        from zero_transformer_b200.utils.metrics import measured_peaks

        pk = measured_peaks()
        fused = getattr(getattr(tr, "engine", None), "_fused", None)
        plane = "nccl (library baseline)" if args.impl != "ours" else ("single GPU" if world == 1 else ("nvls" if getattr(fused, "nvls", False) else "p2p"))
        rec = {
            "metric": f"tokens/sec (whole box, max over ranks) GPT-2 {MODEL_LABEL.get(MODEL, MODEL)} ZeRO-1", "value": round(tok_s, 1), "unit": "tokens/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "cuda_graph": bool(getattr(tr, "use_graph", False)), "ms_per_step": round(ms_dev / args.steps, 3),
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (round(tok_s / base["value"], 3) if base and args.impl == "ours" else None),
            "vs_baseline_source": (base["source"] if base and args.impl == "ours" else None),
            "dtype": "bf16", "data": "synthetic",
            "impl": args.impl,
            "config": {"model": f"gpt2-{MODEL} ALiBi (D={model_cfg.embedding_dim},H={model_cfg.num_head},N={model_cfg.N},V={model_cfg.vocab_size})",
                       "global_batch": ROWS_PER_GPU * world, "seq_len": SEQ, "grad_accum": ACCUM, "dropout": 0.1,
                       "parallelism": f"dp{world}+zero1", "data_plane": plane, "optimizer": "AdamW fp32 master, elementwise clip 1.0",
                       "l2": "working set per step (weights+activations, >30 GB) far exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": round(tok_s_e2e, 1), "unit": "tokens/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": last_loss},
            "gpu_launches": launches,
            "clocks": clocks,
            "mfu_vs_measured_sustained": round(tok_s / world * flops_tok / (pk.get("bf16_tflops_sustained", 1400.0) * 1e12), 4),
            "mfu_dense_attention_flops": round(tok_s / world * flops_tok_dense / (pk.get("bf16_tflops_sustained", 1400.0) * 1e12), 4),
            "peaks_source": pk.get("source"),
        }
        print(json.dumps(rec))
    if hasattr(tr, "close"):
        tr.close()
    pdist.shutdown()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
