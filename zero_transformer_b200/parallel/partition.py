"""Partition rules: which rank owns which slice of every optimizer-state tensor, which bucket it travels in and
whether it is weight-decayed.

The reference expresses ZeRO-1 as an ordered list of regex rules that give every parameter / optimizer-state tensor a
``PartitionSpec`` over the single ``"dp"`` axis (``src/partitioning/partition.py``: ``setup_dp_mesh`` :18-25,
``_match`` / ``_replacement_rules`` :28-87, ``set_partitions_zero`` :90-111, ``create_opt_spec`` :114-140) and asserts that
every sharded dimension divides by the mesh size.  This module is the same decision point for this framework and it is
what ``Zero1Engine`` builds its ``FlatLayout`` from:

  * an ordered rule table (first regex that matches the parameter name wins) yields, per parameter,
      - ``spec``    ``("dp", None, ...)``: the flattened tensor is cut into ``world`` contiguous chunks, rank r owns
                    chunk r (row blocks of a 2-D weight);
      - ``bucket``  the communication bucket (unit of the fused reduce-scatter + AdamW + all-gather launch during
                    backward and of the gather consumed by the forward): one per transformer block, one for the
                    embedding / norms;
      - ``decay``   AdamW weight decay on or off (the reference masks with ``ndim >= 2``, ``main_zero.py:149-156``);
  * no divisibility assertion: regions are padded to ``world * 64`` elements (``layout.py``);
  * ``describe`` / ``owner_of`` / the module CLI print the concrete shard map:

        python -m zero_transformer_b200.parallel.partition --model 1_3b --world 8
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Callable, Dict, List, Mapping, Optional, Sequence, Tuple

from .dist import DistEnv, setup_dp
from .layout import FlatLayout, default_bucket_of

Spec = Tuple[Optional[str], ...]


@dataclass(frozen=True)
class Rule:
    pattern: str
    bucket: Optional[str] = None     # None: ``default_bucket_of(name)``; may use regex groups (``block\\1``)
    decay: Optional[bool] = None     # None: decay iff the tensor has >= 2 dimensions


# First match wins.  The block rule keeps the four GEMM weights of a block in one bucket (they are reduced, updated and
# gathered together as soon as the block's backward has finished); every LayerNorm scale and the tied embedding travel
# in the "embed" bucket, which is gathered before the first block runs.
DEFAULT_RULES: Tuple[Rule, ...] = (
    Rule(r"^wte\.weight$", bucket="embed", decay=True),
    Rule(r"^blocks\.(\d+)\.ln\d\.(weight|bias)$", bucket="embed", decay=False),
    Rule(r"^blocks\.(\d+)\..*\.bias$", bucket=r"block\1", decay=False),
    Rule(r"^blocks\.(\d+)\.", bucket=r"block\1"),
    Rule(r"^ln_f\.", bucket="embed", decay=False),
    Rule(r".*"),
)


def match_rule(name: str, rules: Sequence[Rule] = DEFAULT_RULES) -> Tuple[Rule, "re.Match[str]"]:
    for r in rules:
        m = re.search(r.pattern, name)
        if m:
            return r, m
    raise KeyError(f"no partition rule matches parameter '{name}'")


def setup_dp_mesh(backend: str | None = None) -> DistEnv:
    """1-D data-parallel 'mesh': one process per GPU in a single ``torch.distributed`` group."""
    return setup_dp(backend)


def set_partitions_zero(param_shapes: Mapping[str, Tuple[int, ...]], world: int, rules: Sequence[Rule] = DEFAULT_RULES) -> Dict[str, Spec]:
    """``{param name: spec}``; every parameter must be covered by a rule (as in the reference, an unmatched name is an
    error rather than a silently replicated tensor)."""
    out: Dict[str, Spec] = {}
    for name, shape in param_shapes.items():
        match_rule(name, rules)
        out[name] = ("dp",) + (None,) * (max(len(shape), 1) - 1)
    return out


def create_opt_spec(param_spec: Mapping[str, Spec]) -> Dict[str, Dict[str, Spec]]:
    """Adam ``mu`` / ``nu`` (and the fp32 master copy) inherit the parameter's spec; scalar state (step count) is
    replicated -- the reference's ``create_opt_spec``."""
    return {"master": dict(param_spec), "mu": dict(param_spec), "nu": dict(param_spec), "count": {}}


def bucket_fn(rules: Sequence[Rule] = DEFAULT_RULES) -> Callable[[str], str]:
    def bucket_of(name: str) -> str:
        r, m = match_rule(name, rules)
        return m.expand(r.bucket) if r.bucket is not None else default_bucket_of(name)

    return bucket_of


def decay_mask(named_shapes: Sequence[Tuple[str, Tuple[int, ...]]], rules: Sequence[Rule] = DEFAULT_RULES) -> Dict[str, bool]:
    out = {}
    for name, shape in named_shapes:
        r, _ = match_rule(name, rules)
        out[name] = (len(shape) >= 2) if r.decay is None else bool(r.decay)
    return out


def build_layout(named_shapes: Sequence[Tuple[str, Tuple[int, ...]]], world: int, decay: Optional[Mapping[str, bool]] = None,
                 bucket_of: Optional[Callable[[str], str]] = None, rules: Sequence[Rule] = DEFAULT_RULES) -> FlatLayout:
    """The shard map the engine uses.  ``decay`` / ``bucket_of`` given explicitly (a model's own ``decay_mask()``, a
    test's bucketing) override the rule table."""
    named_shapes = [(n, tuple(s)) for n, s in named_shapes]
    set_partitions_zero(dict(named_shapes), world, rules)     # coverage check
    dm = dict(decay_mask(named_shapes, rules))
    if decay is not None:
        dm.update({n: bool(v) for n, v in decay.items() if n in dm})
    return FlatLayout(named_shapes, dm, world, bucket_of or bucket_fn(rules))


def layout_for(model, world: int, rules: Sequence[Rule] = DEFAULT_RULES) -> FlatLayout:
    """The concrete shard map the engine uses for ``model`` at ``world`` ranks."""
    named = [(n, tuple(p.shape)) for n, p in model.named_parameters() if p.requires_grad]
    decay = model.decay_mask() if hasattr(model, "decay_mask") else None
    return build_layout(named, world, decay, rules=rules)


def owner_of(layout: FlatLayout, name: str, flat_index: int) -> int:
    """Rank that owns element ``flat_index`` (row-major) of parameter ``name``'s optimizer state."""
    rg = next(r for r in layout.regions if r.name == name)
    if not 0 <= flat_index < rg.numel:
        raise IndexError(flat_index)
    return flat_index // rg.chunk


def describe(layout: FlatLayout, dtype_bytes: int = 2) -> List[dict]:
    """One row per parameter: spec, bucket, decay, elements per rank, padding, and rank 0's row range for 2-D weights."""
    rows = []
    for rg in layout.regions:
        cols = 1
        for d in rg.shape[1:]:
            cols *= d
        rows.append({"name": rg.name, "shape": tuple(rg.shape), "spec": ("dp",) + (None,) * (max(len(rg.shape), 1) - 1),
                     "bucket": layout.buckets[rg.bucket].name, "decay": rg.decay, "elems_per_rank": rg.chunk,
                     "pad_elems": rg.padded - rg.numel, "rows_per_rank": round(rg.chunk / max(cols, 1), 2),
                     "state_bytes_per_rank": rg.chunk * 12, "wire_bytes": rg.padded * dtype_bytes})
    return rows


def main(argv=None) -> int:
    import argparse

    from ..config import DEFAULT_MODEL_CFG
    from ..models.gpt import model_getter

    ap = argparse.ArgumentParser(description="print the ZeRO-1 shard map of a model")
    ap.add_argument("--model", default="1_3b")
    ap.add_argument("--model-cfg", default=DEFAULT_MODEL_CFG)
    ap.add_argument("--world", type=int, default=8)
    a = ap.parse_args(argv)
    model = model_getter(a.model, a.model_cfg, device="meta")
    L = layout_for(model, a.world)
    rows = describe(L)
    print(f"{'parameter':42s} {'shape':>16s} {'bucket':>8s} {'decay':>5s} {'elems/rank':>12s} {'pad':>6s}")
    for r in rows:
        print(f"{r['name']:42s} {str(r['shape']):>16s} {r['bucket']:>8s} {str(r['decay']):>5s} {r['elems_per_rank']:12d} {r['pad_elems']:6d}")
    tot = sum(r["elems_per_rank"] for r in rows)
    print(f"world {a.world}: {L.total / 1e6:.1f}M elements in {len(L.buckets)} buckets; per rank {tot / 1e6:.1f}M elements, "
          f"{tot * 12 / 2**30:.2f} GiB of fp32 master + moments (unsharded: {L.total * 12 / 2**30:.2f} GiB)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
