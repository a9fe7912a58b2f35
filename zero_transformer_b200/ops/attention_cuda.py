"""Autograd wrapper of the tcgen05 ALiBi flash-attention kernels (csrc/attention.cu).

Same math as the reference's ``CausalAttention`` core (``src/models/layers.py:145-178``: scaled scores + ALiBi
bias ``slope_h * (j - i)`` from ``get_slopes``/``create_mask`` ``layers.py:17-44``, causal mask, fp32 softmax,
dropout on the probabilities, ``P V``) without materialising the ``(b, nh, T, T)`` score tensor."""
from __future__ import annotations

import torch

from . import cuda as K


def supported(qkv: torch.Tensor, n_head: int) -> bool:
    B, T, D3 = qkv.shape
    hd = D3 // 3 // n_head
    return qkv.dtype == torch.bfloat16 and T % 128 == 0 and hd in (64, 96, 128)


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, slopes, n_head, scale, p):
        qkv = qkv.contiguous()
        rng = K.rng_state(qkv.device)
        op_id = K._next_op_id() if p > 0 else 0
        sl = slopes.float().contiguous() if slopes is not None else None
        out, lse = K._C.attention_fwd(qkv, sl, n_head, scale, p, rng if p > 0 else None, op_id)
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (sl, n_head, scale, p, op_id)
        K._note_attention_output(out, n_head, qkv.shape[1])
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        sl, n_head, scale, p, op_id = ctx.meta
        rng = K.rng_state(qkv.device)
        dout = dout.contiguous()
        delta = K._DELTA_STASH.pop(dout.data_ptr(), None)     # rowsum(dO o O) from the out-projection dgrad GEMM, if it ran fused
        if delta is not None and tuple(delta.shape) != (qkv.shape[0], n_head, qkv.shape[1]):
            delta = None
        dqkv = K._C.attention_bwd(qkv, out, dout, lse, sl, n_head, scale, p, rng if p > 0 else None, op_id,
                                  dq_bf16=_dq_bf16(qkv.shape[1]), delta=delta)
        return dqkv, None, None, None, None


def _dq_bf16(T: int) -> bool:
    mode = K.CONFIG.attn_dq_bf16
    if isinstance(mode, bool):
        return mode
    if mode == "auto":
        return T <= 2048 and not K.CONFIG.deterministic
    return mode == "1"


def alibi_attention(qkv, slopes, n_head, scale, p):
    return _Attention.apply(qkv, slopes, n_head, scale, float(p))
