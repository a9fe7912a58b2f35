"""API-surface parity with the reference's partitioning helpers
(``src/partitioning/partition.py``: ``setup_dp_mesh`` :18-25, ``set_partitions_zero`` :90-111,
``create_opt_spec`` :114-140).  The reference expresses ZeRO-1 as regex rules that give every
optimizer-state tensor a ``PartitionSpec`` over the single ``"dp"`` axis; here the same decision is the
``FlatLayout`` shard map (every parameter split into ``world`` contiguous chunks = ``P("dp", None)`` for
a 2-D weight, ``P("dp")`` for a LayerNorm scale).  These wrappers return that map in a form a user of
the reference will recognise.
"""
from __future__ import annotations

from typing import Dict, Mapping, Tuple

from .dist import DistEnv, setup_dp
from .layout import FlatLayout


def setup_dp_mesh(backend: str | None = None) -> DistEnv:
    """1-D data-parallel 'mesh': one process per GPU in a single ``torch.distributed`` group."""
    return setup_dp(backend)


def set_partitions_zero(param_shapes: Mapping[str, Tuple[int, ...]], world: int) -> Dict[str, Tuple[str, ...]]:
    """``{param name: spec}`` with spec ``("dp", None, ...)``: dim 0 of every tensor is split over the DP
    ranks (contiguous row chunks).  Unlike the reference there is no divisibility assertion: regions are
    padded to ``world * 64`` elements."""
    out = {}
    for name, shape in param_shapes.items():
        out[name] = ("dp",) + (None,) * (len(shape) - 1)
    return out


def create_opt_spec(param_spec: Mapping[str, Tuple[str, ...]]) -> Dict[str, Dict[str, Tuple[str, ...]]]:
    """Adam ``mu`` / ``nu`` (and the fp32 master copy) inherit the parameter's spec; scalar state
    (step count) is replicated -- the reference's ``create_opt_spec``."""
    return {"master": dict(param_spec), "mu": dict(param_spec), "nu": dict(param_spec), "count": {}}


def layout_for(model, world: int) -> FlatLayout:
    """The concrete shard map the engine uses for ``model`` at ``world`` ranks."""
    named = [(n, tuple(p.shape)) for n, p in model.named_parameters() if p.requires_grad]
    decay = model.decay_mask() if hasattr(model, "decay_mask") else {}
    return FlatLayout(named, decay, world)
