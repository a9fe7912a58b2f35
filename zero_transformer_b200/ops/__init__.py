"""Op dispatch: CUDA tensors go to the hand-written sm_100a kernels
(``ops/cuda.py`` -> ``csrc/``), everything else to the PyTorch reference
implementations in ``ops/reference.py``.

There is deliberately no silent fallback on a GPU: if a tensor is on a CUDA
device and the extension is missing, importing ``ops.cuda`` raises.
Set ``ZTB_FORCE_REFERENCE=1`` to run the reference ops on a GPU (used only by
the numerics tests and the NCCL baseline arm, never by the product path).
"""
from __future__ import annotations

import os

import torch

from . import reference
from .reference import alibi_bias, default_attn_scale, get_slopes  # noqa: F401

_FORCE_REF = os.environ.get("ZTB_FORCE_REFERENCE", "0") == "1"


def force_reference(flag: bool) -> None:
    global _FORCE_REF
    _FORCE_REF = bool(flag)


def _use_cuda(t: torch.Tensor) -> bool:
    return t.is_cuda and not _FORCE_REF


def _k():
    from . import cuda  # raises loudly if the extension is not built

    return cuda


def layernorm(x, weight, eps):
    if _use_cuda(x):
        return _k().layernorm(x, weight, eps)
    return reference.layernorm(x, weight, eps)


def layernorm_res(x, weight, eps):
    """``(layernorm(x), x)``: the second output is ``x`` itself, to be used for the residual add, so
    that on CUDA the residual gradient is folded into the LayerNorm backward kernel."""
    if _use_cuda(x):
        return _k().layernorm_res(x, weight, eps)
    return reference.layernorm(x, weight, eps), x


def linear(x, w):
    if _use_cuda(x):
        return _k().linear(x, w)
    return reference.linear(x, w)


def linear_gelu(x, w, approximate: str = "tanh"):
    if _use_cuda(x) and approximate == "tanh":
        return _k().linear_gelu(x, w)
    return reference.linear_gelu(x, w, approximate)


def linear_dropout_add(x, w, residual, p: float, training: bool):
    if _use_cuda(x):
        return _k().linear_dropout_add(x, w, residual, p, training)
    return reference.linear_dropout_add(x, w, residual, p, training)


def mlp(x, w1, w2, residual, p: float, training: bool, approximate: str = "tanh"):
    """``residual + dropout(gelu(x @ w1.T) @ w2.T)`` (reference MLPBlock, ``layers.py:47-77``)."""
    if _use_cuda(x) and approximate == "tanh":
        return _k().mlp(x, w1, w2, residual, p, training)
    return reference.linear_dropout_add(reference.linear_gelu(x, w1, approximate), w2, residual, p, training)


def begin_forward(device) -> None:
    """Start a new dropout stream (one per training forward / micro-step)."""
    import torch as _t

    if _t.device(device).type == "cuda" and not _FORCE_REF:
        _k().begin_forward(device)


def dropout_add(y, residual, p: float, training: bool):
    if _use_cuda(y):
        return _k().dropout_add(y, residual, p, training)
    return reference.dropout_add(y, residual, p, training)


def attention(qkv, slopes, n_head, scale, p=0.0, training=False, kv_cache=None):
    if _use_cuda(qkv) and kv_cache is None:
        return _k().attention(qkv, slopes, n_head, scale, p, training)
    return reference.attention(qkv, slopes, n_head, scale, p, training, kv_cache)


def embedding(idx, wte):
    if _use_cuda(wte):
        return _k().embedding(idx, wte)
    return reference.embedding(idx, wte)


def cross_entropy(logits, labels):
    if _use_cuda(logits):
        return _k().cross_entropy(logits, labels)
    return reference.cross_entropy(logits, labels)


def lm_head_loss(h, wte, tokens):
    if _use_cuda(h):
        return _k().lm_head_loss(h, wte, tokens)
    return reference.lm_head_loss(h, wte, tokens)
