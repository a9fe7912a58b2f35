"""Process bootstrap: one process per GPU, ``torch.distributed`` for the
control plane (NCCL on GPUs, gloo on CPU).  The reference's equivalent is
implicit in ``jax[tpu]`` on a TPU-VM pod plus the 1-D ``Mesh(devices, ("dp",))``
of ``main_zero.py:227-228`` / ``src/partitioning/partition.py:18-25``.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


@dataclass
class DistEnv:
    rank: int
    local_rank: int
    world: int
    device: torch.device
    backend: Optional[str]

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def setup_dp(backend: Optional[str] = None, device: Optional[str] = None, timeout_s: int = 600) -> DistEnv:
    """Initialise the data-parallel group from torchrun-style env vars
    (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  With no env
    vars this is a single-process run and no process group is created."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    use_cuda = torch.cuda.is_available() and device != "cpu"
    if use_cuda:
        dev = torch.device("cuda", local_rank % max(torch.cuda.device_count(), 1))
        torch.cuda.set_device(dev)
    else:
        dev = torch.device("cpu")
    be = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        be = backend or ("nccl" if use_cuda else "gloo")
        if not dist.is_initialized():
            kw = dict(backend=be, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=timeout_s))
            if use_cuda:
                kw["device_id"] = dev
            dist.init_process_group(**kw)
    return DistEnv(rank, local_rank, world, dev, be)


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def all_reduce_mean(x: torch.Tensor) -> torch.Tensor:
    """Mean over the DP group (reference ``pmean`` of the loss,
    ``xmap_train_functions.py:83,103``)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
        x = x / dist.get_world_size()
    return x


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if torch.device(device).type == "cuda" else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
