"""Logging, metrics sinks and step timing.

Reference: process-0 ``logging`` at DEBUG, W&B run with the
flattened config and per-step ``train/loss``, ``train/ppl``, ``Train Sequence
Length``, ``Learning Rate``, ``Tokens Seen (B)``, ``validation/*``
(``main_zero.py:354-366,504-529,561-562,587-588``).  The reference has no
timers at all; here every step is timed with CUDA events (max over
ranks on demand) and reported as tokens/s and MFU against the measured peak.
"""
from __future__ import annotations

import json
import logging
import os
import time
from typing import Any, Dict, Optional

import torch

_PEAKS_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measured_peaks() -> Dict[str, Any]:
    """Roofline denominators: driver-measured file, else the profiling guide's fallback."""
    try:
        with open(_PEAKS_PATH) as f:
            d = json.load(f)
        d["source"] = "measured"
        return d
    except (OSError, ValueError):
        return dict(FALLBACK_PEAKS)


def get_logger(name: str = "ztb", is_main: bool = True) -> logging.Logger:
    log = logging.getLogger(name)
    if not log.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(message)s"))
        log.addHandler(h)
    log.setLevel(logging.DEBUG if is_main else logging.ERROR)
    log.propagate = False
    return log


def compute_tokens_seen(step: int, global_rows: int, seq_len: int) -> int:
    """Exact tokens consumed after ``step`` optimizer updates (the reference's
    ``compute_tokens_seen``, ``src/training/training_utils.py:32-34``, over-counts
    by ``num_host``)."""
    return int(step) * int(global_rows) * int(seq_len)


class MetricsSink:
    """JSONL file always; Weights & Biases additionally when a project is
    configured and ``wandb`` can start (offline mode -- there is no network)."""

    def __init__(self, out_dir: str, run_config: Dict[str, Any], project: str = "", run_id: Optional[str] = None, enabled: bool = True):
        self.enabled = enabled
        self.wandb = None
        self.f = None
        if not enabled:
            return
        os.makedirs(out_dir, exist_ok=True)
        self.f = open(os.path.join(out_dir, "metrics.jsonl"), "a", buffering=1)
        self.f.write(json.dumps({"_config": run_config}, default=str) + "\n")
        if project:
            try:
                import wandb

                os.environ.setdefault("WANDB_MODE", "offline")
                self.wandb = wandb.init(project=project, id=run_id, resume="allow", config=run_config, dir=out_dir)
            except Exception as e:  # pragma: no cover - depends on the wandb install
                print(f"[metrics] wandb disabled: {e}")
                self.wandb = None

    def log(self, metrics: Dict[str, Any], step: int) -> None:
        if not self.enabled:
            return
        rec = {"step": step, "time": time.time()}
        rec.update({k: (float(v) if isinstance(v, (int, float)) or hasattr(v, "item") else v) for k, v in metrics.items()})
        self.f.write(json.dumps(rec, default=str) + "\n")
        if self.wandb is not None:
            self.wandb.log(metrics, step=step)

    def close(self) -> None:
        if self.f:
            self.f.close()
        if self.wandb is not None:
            self.wandb.finish()


class StepTimer:
    """CUDA-event timer on the current stream (wall clock on CPU)."""

    def __init__(self, device):
        self.cuda = torch.device(device).type == "cuda"
        if self.cuda:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
        self.t0 = 0.0

    def start(self):
        if self.cuda:
            self.e0.record()
        else:
            self.t0 = time.perf_counter()

    def stop_ms(self) -> float:
        if self.cuda:
            self.e1.record()
            self.e1.synchronize()
            return self.e0.elapsed_time(self.e1)
        return (time.perf_counter() - self.t0) * 1e3


def mfu(tokens_per_s_per_gpu: float, flops_per_token: float, sustained: bool = True) -> float:
    pk = measured_peaks()
    peak = pk.get("bf16_tflops_sustained" if sustained else "bf16_tflops", pk["bf16_tflops"]) * 1e12
    return tokens_per_s_per_gpu * flops_per_token / peak
