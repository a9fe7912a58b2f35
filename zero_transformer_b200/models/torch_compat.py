"""Inference-side loader with the reference's signature:
``model_getter(model_size, config_path, model_checkpoint)`` (``torch_compatability/GPT2.py:448-474``).
Accepts the reference's torch-side table spelling (``num_ctx``) and a reference-format ``.pth``; returns
our ``Transformer`` (same math, KV cache included) with the weights loaded strictly."""
from __future__ import annotations

from typing import Optional

import torch

from ..config import DEFAULT_MODEL_CFG
from ..utils.export import from_reference_state_dict
from .gpt import Transformer, model_getter as _model_getter


def model_getter(model_size: str, config_path: str = DEFAULT_MODEL_CFG, model_checkpoint: Optional[str] = None) -> Transformer:
    # the torch consumer uses erf GELU and LayerNorm eps 1e-5 (GPT2.py:66, nn.LayerNorm default)
    model = _model_getter(model_size, config_path, dtype=torch.float32, gelu="erf", dropout=0.0)
    if model_checkpoint is not None:
        sd = torch.load(model_checkpoint, map_location="cpu", weights_only=False)
        model.load_state_dict(from_reference_state_dict(sd), strict=True)
    return model.eval()
