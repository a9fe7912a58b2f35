"""Checkpoint / resume / warm-start.

Reference behaviour (``main_zero.py:58-139,268-313,326-342,554-575``): separate
``params_<step>`` and ``optimizer_<step>`` checkpoints under
``<bucket>/<checkpoint_directory>/{params,optimizer}``, ``keep=5``, written by
process 0 only after gathering the sharded optimizer state, fp32 on disk
(``logs/580.md:100-106``), restore rebuilds the optimizer state and adopts the
step; a fresh (non-resume) start wipes old checkpoints.

Here: the same directory layout on the local filesystem, ``torch.save`` files
(``params_<step>.pt`` = ``{"step", "params": {name: fp32 tensor}}``;
``optimizer_<step>.pt`` = ``{"step", "exp_avg", "exp_avg_sq"}``), world-size
independent because they hold consolidated per-parameter tensors.  A faster
per-rank sharded format (no gather) is available with ``sharded=True``.
"""
from __future__ import annotations

import glob
import os
import re
import shutil
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def _dirs(workdir: str) -> Tuple[str, str]:
    return os.path.join(workdir, "params"), os.path.join(workdir, "optimizer")


def _steps(d: str, prefix: str) -> List[int]:
    out = []
    for f in glob.glob(os.path.join(d, f"{prefix}*")):
        m = re.match(rf"{prefix}(\d+)(\.pt|\.rank\d+\.pt)?$", os.path.basename(f))
        if m:
            out.append(int(m.group(1)))
    return sorted(set(out))


def _prune(d: str, prefix: str, keep: int) -> None:
    steps = _steps(d, prefix)
    for s in steps[:-keep] if keep > 0 else []:
        for f in glob.glob(os.path.join(d, f"{prefix}{s}.*")) + glob.glob(os.path.join(d, f"{prefix}{s}")):
            try:
                os.remove(f)
            except OSError:
                pass


def _atomic_save(obj, path: str) -> None:
    tmp = path + ".tmp"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def wipe_checkpoints(workdir: str) -> None:
    """Fresh start removes old checkpoints (reference ``main_zero.py:326-342``)."""
    for d in _dirs(workdir):
        if os.path.isdir(d):
            shutil.rmtree(d)


def _rng_snapshot(engine):
    """(seed, offset) of the device-resident Philox state that keys every dropout mask: restored on resume so the
    continued run draws the masks the uninterrupted run would have drawn (the reference folds resume_step into its
    key instead, ``main_zero.py:432``)."""
    if engine.device.type != "cuda":
        return None
    from ..ops import cuda as K

    return [int(v) for v in K.rng_state(engine.device).tolist()]


def _rng_restore(engine, snap) -> None:
    if snap is None or engine.device.type != "cuda":
        return
    from ..ops import cuda as K

    K.rng_state(engine.device).copy_(torch.tensor([int(snap[0]), int(snap[1])], dtype=torch.int64))


def stable_run_id(workdir: str) -> str:
    """One metrics-run id per checkpoint directory, created on first use and re-read on resume, so a resumed job
    continues the same W&B run (the reference generates a fresh id on every start, ``main_zero.py:355-356``)."""
    path = os.path.join(workdir, "run_id.txt")
    try:
        with open(path) as f:
            rid = f.read().strip()
        if rid:
            return rid
    except OSError:
        pass
    import uuid

    rid = uuid.uuid4().hex[:8]
    os.makedirs(workdir, exist_ok=True)
    with open(path, "w") as f:
        f.write(rid)
    return rid


def save_checkpoint(engine, workdir: str, step: int, keep: int = 5, sharded: bool = False) -> None:
    """Collective: every rank must call.  Rank 0 writes the consolidated files."""
    pdir, odir = _dirs(workdir)
    rank = engine.rank
    rng = _rng_snapshot(engine)
    if rank == 0:
        os.makedirs(pdir, exist_ok=True)
        os.makedirs(odir, exist_ok=True)
    if dist.is_initialized():
        dist.barrier()
    if sharded:
        _atomic_save(engine.shard_state(), os.path.join(odir, f"optimizer_{step}.rank{rank}.pt"))
        params = engine.full_params_fp32(0)
        if rank == 0:
            _atomic_save({"step": step, "params": params, "rng": rng, "opt_step": engine.step_count}, os.path.join(pdir, f"params_{step}.pt"))
    else:
        st = engine.consolidated_state(0)
        if rank == 0:
            _atomic_save({"step": step, "params": st["params"], "rng": rng, "opt_step": engine.step_count}, os.path.join(pdir, f"params_{step}.pt"))
            _atomic_save({"step": step, "opt_step": engine.step_count, "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]},
                         os.path.join(odir, f"optimizer_{step}.pt"))
    if rank == 0:
        _prune(pdir, "params_", keep)
        _prune(odir, "optimizer_", keep)
    if dist.is_initialized():
        dist.barrier()


def save_checkpoint_params(engine, workdir: str, step: int, keep: int = 5) -> None:
    """Reference name (``main_zero.py:58-71``): parameters only.  Collective."""
    pdir = _dirs(workdir)[0]
    params = engine.full_params_fp32(0)
    if engine.rank == 0:
        os.makedirs(pdir, exist_ok=True)
        _atomic_save({"step": step, "params": params}, os.path.join(pdir, f"params_{step}.pt"))
        _prune(pdir, "params_", keep)


def save_checkpoint_optimizer(engine, workdir: str, step: int, keep: int = 5) -> None:
    """Reference name (``main_zero.py:74-93``): gathered Adam moments only.  Collective."""
    odir = _dirs(workdir)[1]
    st = engine.consolidated_state(0)
    if engine.rank == 0:
        os.makedirs(odir, exist_ok=True)
        _atomic_save({"step": step, "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]}, os.path.join(odir, f"optimizer_{step}.pt"))
        _prune(odir, "optimizer_", keep)


def latest_step(workdir: str) -> Optional[int]:
    steps = _steps(_dirs(workdir)[0], "params_")
    return steps[-1] if steps else None


def restore_param_checkpoint(workdir: str, step: Optional[int] = None) -> Tuple[Dict[str, torch.Tensor], int]:
    """-> (params dict, step)   (reference ``restore_param_checkpoint``, ``main_zero.py:96-102``)."""
    pdir = _dirs(workdir)[0]
    step = latest_step(workdir) if step is None else step
    if step is None:
        raise FileNotFoundError(f"no params_<step>.pt under {pdir}")
    st = torch.load(os.path.join(pdir, f"params_{step}.pt"), map_location="cpu", weights_only=False)
    return st["params"], int(st["step"])


def restore_opt_checkpoint(workdir: str, step: Optional[int] = None, rank: int = 0):
    """-> (state dict with exp_avg/exp_avg_sq [or a shard state], step)
    (reference ``restore_opt_checkpoint``, ``main_zero.py:105-139``)."""
    odir = _dirs(workdir)[1]
    if step is None:
        steps = _steps(odir, "optimizer_")
        if not steps:
            raise FileNotFoundError(f"no optimizer_<step> under {odir}")
        step = steps[-1]
    full = os.path.join(odir, f"optimizer_{step}.pt")
    if os.path.exists(full):
        st = torch.load(full, map_location="cpu", weights_only=False)
        return st, int(st["step"])
    shard = os.path.join(odir, f"optimizer_{step}.rank{rank}.pt")
    st = torch.load(shard, map_location="cpu", weights_only=False)
    return st, int(st["step"])


def resume(engine, workdir: str, step: Optional[int] = None) -> int:
    """Load params + optimizer state (+ the dropout RNG position) into the engine and adopt the step.
    Returns the training step to continue from; the optimizer's own counter (Adam bias correction and LR schedule)
    is restored separately -- after a warm start the two differ."""
    pdir = _dirs(workdir)[0]
    pstep = latest_step(workdir) if step is None else step
    if pstep is None:
        raise FileNotFoundError(f"no params_<step>.pt under {pdir}")
    pst = torch.load(os.path.join(pdir, f"params_{pstep}.pt"), map_location="cpu", weights_only=False)
    params = pst["params"]
    ost, ostep = restore_opt_checkpoint(workdir, pstep, engine.rank)
    if "master" in ost:  # per-rank shard format
        engine.load_shard_state(ost)
    else:
        engine.load_consolidated_state({"step": int(ost.get("opt_step", ostep)), "params": params, "exp_avg": ost["exp_avg"], "exp_avg_sq": ost["exp_avg_sq"]})
    _rng_restore(engine, pst.get("rng"))
    return int(pstep)


def warm_start(engine, warm_dir: str, extend_depth: bool = False) -> None:
    """Initialise weights -- and, when the other run's optimizer checkpoint is there, the Adam moments together with
    the optimizer's update counter -- from another run's checkpoint (reference ``cfg.model.warm_init``,
    ``main_zero.py:268-289``); optional Gopher-style depth extension.

    The *training* step stays 0 (data position, logging, total_steps), exactly as in the reference, but the optimizer
    counter is adopted with the moments: the reference restores ``ScaleByAdamState.count`` and the schedule count next to
    ``mu`` / ``nu`` (``main_zero.py:105-139``).  Loading warm moments under a count of 0 would divide them by
    ``1 - beta^t`` with t = 1, 2, ... and inflate the first updates ~2x."""
    from .extend_params import extend_params

    params, _ = restore_param_checkpoint(warm_dir)
    n_target = len({n.split(".")[1] for n in engine.layout.by_name if n.startswith("blocks.")})
    st = {"step": 0, "params": params}
    load_moments = False
    try:
        ost, ostep = restore_opt_checkpoint(warm_dir)
        if "exp_avg" in ost:
            st["exp_avg"], st["exp_avg_sq"] = ost["exp_avg"], ost["exp_avg_sq"]
            st["step"] = int(ost.get("opt_step", ostep))
            load_moments = True
    except FileNotFoundError:
        pass
    if extend_depth:
        st["params"] = extend_params(st["params"], n_target)
        if load_moments:
            st["exp_avg"] = extend_params(st["exp_avg"], n_target)
            st["exp_avg_sq"] = extend_params(st["exp_avg_sq"], n_target)
    engine.load_consolidated_state(st, load_moments=load_moments, adopt_step=load_moments)
