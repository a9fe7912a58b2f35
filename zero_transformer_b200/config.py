"""YAML config system.

Same two-file surface as the reference (``conf/config.yaml`` with
``training/model/data`` groups -- reference ``conf/config.yaml:1-29`` -- and a
named model table ``conf/model_config.yaml:1-35``) but without OmegaConf
(not installed here): a tiny attribute-dict over PyYAML plus ``a.b.c=value``
command-line overrides, which the reference does not have.
"""
from __future__ import annotations

import copy
import os
from typing import Any, Dict, Iterable, Iterator, Mapping

import yaml

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_CFG = os.path.join(_REPO_ROOT, "conf", "config.yaml")
DEFAULT_MODEL_CFG = os.path.join(_REPO_ROOT, "conf", "model_config.yaml")


class Config(Mapping):
    """Nested read/write attribute access over a plain dict.

    ``cfg.training.batch_size`` and ``cfg["training"]["batch_size"]`` both
    work; unknown keys raise ``AttributeError``/``KeyError`` instead of
    silently returning ``None``.
    """

    def __init__(self, data: Mapping[str, Any] | None = None):
        object.__setattr__(self, "_d", {})
        for k, v in (data or {}).items():
            self._d[k] = _wrap(v)

    # mapping protocol -------------------------------------------------
    def __getitem__(self, k: str) -> Any:
        return self._d[k]

    def __setitem__(self, k: str, v: Any) -> None:
        self._d[k] = _wrap(v)

    def __iter__(self) -> Iterator[str]:
        return iter(self._d)

    def __len__(self) -> int:
        return len(self._d)

    def __contains__(self, k: object) -> bool:
        return k in self._d

    # attribute protocol -----------------------------------------------
    def __getattr__(self, k: str) -> Any:
        try:
            return self._d[k]
        except KeyError as e:
            raise AttributeError(f"config has no key '{k}' (have {list(self._d)})") from e

    def __setattr__(self, k: str, v: Any) -> None:
        self._d[k] = _wrap(v)

    def get(self, k: str, default: Any = None) -> Any:
        return self._d.get(k, default)

    def to_dict(self) -> Dict[str, Any]:
        return {k: (v.to_dict() if isinstance(v, Config) else copy.deepcopy(v)) for k, v in self._d.items()}

    def __repr__(self) -> str:
        return f"Config({self.to_dict()!r})"

    def __deepcopy__(self, memo):
        return Config(self.to_dict())

    # overrides --------------------------------------------------------
    def set_path(self, dotted: str, value: Any) -> None:
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node:
                node[p] = {}
            node = node[p]
        node[parts[-1]] = value

    def merge_overrides(self, overrides: Iterable[str]) -> "Config":
        """Apply ``a.b=value`` strings (value parsed as YAML)."""
        for ov in overrides:
            if "=" not in ov:
                raise ValueError(f"override '{ov}' is not of the form key=value")
            k, v = ov.split("=", 1)
            self.set_path(k.strip(), yaml.safe_load(v))
        return self


def _wrap(v: Any) -> Any:
    if isinstance(v, Config):
        return v
    if isinstance(v, Mapping):
        return Config(v)
    return v


def load_yaml(path: str) -> Config:
    with open(path, "r") as f:
        data = yaml.safe_load(f) or {}
    if not isinstance(data, Mapping):
        raise ValueError(f"{path}: top level of a config file must be a mapping")
    return Config(data)


def load_config(path: str | None = None, overrides: Iterable[str] = ()) -> Config:
    cfg = load_yaml(path or DEFAULT_CFG)
    cfg.merge_overrides(overrides)
    validate_train_config(cfg)
    return cfg


_REQUIRED = {
    "training": [
        "batch_size", "peak_learning_rate", "warmup_steps", "total_steps",
        "end_learning_rate", "weight_decay", "gradient_accumulation_steps",
        "evaluation_frequency", "maximum_evaluation_steps", "train_context",
    ],
    "model": ["size"],
    "data": ["max_context", "checkpoint_directory"],
}


def validate_train_config(cfg: Config) -> None:
    """The reference validates nothing beyond two asserts; we
    check presence and basic arithmetic so a bad YAML fails before a 1.3B
    model is built."""
    for group, keys in _REQUIRED.items():
        if group not in cfg:
            raise ValueError(f"config is missing the '{group}' group")
        for k in keys:
            if k not in cfg[group]:
                raise ValueError(f"config is missing '{group}.{k}'")
    t, d = cfg.training, cfg.data
    if d.max_context % t.train_context != 0:
        raise ValueError("data.max_context must be a multiple of training.train_context")
    if t.gradient_accumulation_steps < 1:
        raise ValueError("training.gradient_accumulation_steps must be >= 1")


def flatten_dict(d: Mapping[str, Any], parent_key: str = "", sep: str = ".") -> Dict[str, Any]:
    """Nested mapping -> ``{"a.b": v}``; used for the run-config logged to
    W&B/JSONL (capability of reference ``src/utils/configs.py:7-17``)."""
    out: Dict[str, Any] = {}
    stack = [(parent_key, d)]
    while stack:
        prefix, node = stack.pop()
        for k, v in node.items():
            key = f"{prefix}{sep}{k}" if prefix else str(k)
            if isinstance(v, Mapping):
                stack.append((key, v))
            else:
                out[key] = v
    return dict(sorted(out.items()))
