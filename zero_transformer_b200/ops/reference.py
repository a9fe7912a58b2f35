"""Plain-PyTorch implementations of every op the model uses.

These are (a) the CPU path (gloo plumbing tests, BASELINE config #1) and
(b) the fp32 oracles the CUDA kernels are tested against.  Semantics follow
the reference model (``src/models/layers.py:47-191``, ``src/models/GPT.py:68-113``,
``src/utils/losses.py:9-23``) with the documented decisions of SURVEY 7.1.
All functions are differentiable through ordinary autograd.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn.functional as F


def get_slopes(n: int) -> List[float]:
    """ALiBi head slopes (geometric sequence; non-power-of-two head counts
    interleave the next power of two).  Capability of reference
    ``src/models/layers.py:17-30`` / ``torch_compatability/GPT2.py:135-150``."""
    def pow2(m: int) -> List[float]:
        base = 2.0 ** (-8.0 / m)
        return [base ** (i + 1) for i in range(m)]

    if n & (n - 1) == 0:
        return pow2(n)
    lo = 1 << (n.bit_length() - 1)
    return pow2(lo) + get_slopes(2 * lo)[0::2][: n - lo]


def alibi_bias(slopes: torch.Tensor, t_q: int, t_k: int) -> torch.Tensor:
    """Canonical ``slope_h * (j - i)`` bias, (H, t_q, t_k), queries aligned to
    the *end* of the keys (so t_q=1, t_k=n is the KV-cache decode row).  The
    reference's "last row" form (``layers.py:33-44``) differs by a per-row
    constant, which softmax removes."""
    i = torch.arange(t_k - t_q, t_k, device=slopes.device).view(1, t_q, 1)
    j = torch.arange(t_k, device=slopes.device).view(1, 1, t_k)
    return slopes.view(-1, 1, 1).float() * (j - i).float()


def layernorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """Bias-free LayerNorm, statistics in fp32 (reference ``GPT.py:42,49,98``)."""
    xf = x.float()
    mu = xf.mean(-1, keepdim=True)
    var = (xf - mu).pow(2).mean(-1, keepdim=True)
    return ((xf - mu) * torch.rsqrt(var + eps) * weight.float()).to(x.dtype)


def gelu(x: torch.Tensor, approximate: str = "tanh") -> torch.Tensor:
    return F.gelu(x, approximate="tanh" if approximate == "tanh" else "none")


def linear(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``x @ w.T`` with ``w`` stored (out, in) -- K-major for both GEMM
    operands, which is what the tcgen05 kernel consumes without a transpose."""
    return F.linear(x, w.to(x.dtype))


def linear_gelu(x: torch.Tensor, w: torch.Tensor, approximate: str = "tanh") -> torch.Tensor:
    return gelu(linear(x, w), approximate)


def dropout_add(y: torch.Tensor, residual: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """``residual + dropout(y)`` (reference ``layers.py:77,191`` + ``GPT.py:43-49``)."""
    return residual + F.dropout(y, p, training)


def linear_dropout_add(x, w, residual, p: float, training: bool) -> torch.Tensor:
    return dropout_add(linear(x, w), residual, p, training)


def attention(
    qkv: torch.Tensor,
    slopes: Optional[torch.Tensor],
    n_head: int,
    scale: float,
    p: float = 0.0,
    training: bool = False,
    kv_cache: Optional[tuple] = None,
) -> torch.Tensor:
    """Causal (ALiBi-biased) attention over a packed ``(B, T, 3*D)`` tensor
    laid out ``[q | k | v]`` along the last dim, heads contiguous inside each.
    Scores and softmax in fp32 (reference ``layers.py:159-175``).  With
    ``kv_cache=(k_past, v_past)`` (each ``(B, H, T_past, hd)``) the new keys are
    appended (reference ``GPT2.py:176-182``) and ``(out, (k, v))`` is returned.
    """
    B, T, D3 = qkv.shape
    D = D3 // 3
    hd = D // n_head
    q, k, v = qkv.view(B, T, 3, n_head, hd).permute(2, 0, 3, 1, 4)  # each (B,H,T,hd)
    if kv_cache is not None and kv_cache[0] is not None:
        k = torch.cat([kv_cache[0], k], dim=2)
        v = torch.cat([kv_cache[1], v], dim=2)
    Tk = k.shape[2]
    if qkv.is_cuda and kv_cache is None:
        # stock-PyTorch GPU path (baseline arm): SDPA with the ALiBi+causal additive mask
        i = torch.arange(T, device=qkv.device).view(T, 1)
        j = torch.arange(T, device=qkv.device).view(1, T)
        bias = alibi_bias(slopes, T, T) if slopes is not None else torch.zeros(1, T, T, device=qkv.device)
        bias = bias.masked_fill(j > i, float("-inf")).to(q.dtype).unsqueeze(0)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, dropout_p=p if training else 0.0, scale=scale)
        return o.transpose(1, 2).reshape(B, T, D)
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if slopes is not None:
        s = s + alibi_bias(slopes, T, Tk).unsqueeze(0)
    i = torch.arange(Tk - T, Tk, device=qkv.device).view(T, 1)
    j = torch.arange(Tk, device=qkv.device).view(1, Tk)
    s = s.masked_fill(j > i, torch.finfo(torch.float32).min)
    pr = torch.softmax(s, dim=-1)
    pr = F.dropout(pr, p, training)
    o = torch.matmul(pr.to(v.dtype), v)  # (B,H,T,hd)
    o = o.transpose(1, 2).reshape(B, T, D)
    if kv_cache is not None:
        return o, (k, v)
    return o


def embedding(idx: torch.Tensor, wte: torch.Tensor) -> torch.Tensor:
    return F.embedding(idx, wte)


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Mean over all rows of ``-log_softmax(fp32 logits)[label]``
    (reference ``src/utils/losses.py:21-23``; always fp32)."""
    lf = logits.float()
    return F.cross_entropy(lf.reshape(-1, lf.shape[-1]), labels.reshape(-1).long())


def lm_head_loss(h: torch.Tensor, wte: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
    """Tied LM head + next-token loss: position t predicts ``tokens[t+1]``,
    mean over ``B*(T-1)`` (reference ``GPT.py:100-111``)."""
    logits = linear(h[:, :-1], wte)
    return cross_entropy(logits, tokens[:, 1:])


def adamw_step_(p, g, m, v, *, lr, beta1, beta2, eps, wd, step, clip, grad_scale=1.0):
    """One AdamW update on flat fp32 tensors, in place.  ``clamp(g*scale, +-clip)``
    -> Adam moments -> bias correction -> ``p -= lr*(mhat/(sqrt(vhat)+eps) + wd*p)``.
    This is the optax chain of reference ``main_zero.py:160-168``
    (``clip(1.0)`` is *elementwise*).  ``step`` counts from 1."""
    g = g.float() * grad_scale
    if clip is not None and clip > 0:
        g = g.clamp(-clip, clip)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    upd = (m / bc1) / ((v / bc2).sqrt() + eps)
    if isinstance(wd, torch.Tensor):
        upd = upd + wd * p
    elif wd:
        upd = upd + wd * p
    p.sub_(upd * lr)
    return p


def default_attn_scale(hd: int, seq_len: int, mode: str) -> float:
    """``head_dim`` (standard; what the exported torch model uses) or
    ``seq_len`` (the reference's flax quirk, ``layers.py:155,159-161``)."""
    return 1.0 / math.sqrt(hd if mode == "head_dim" else seq_len)
