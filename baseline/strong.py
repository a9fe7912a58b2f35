"""Reference-equivalent baseline arm: the SAME algorithm (ALiBi GPT-2, bf16 compute / fp32 master, ZeRO-1 AdamW with
elementwise clip, dropout 0.1, gradient accumulation) built the way a competent engineer builds it from stock libraries:

  * GEMMs               ``F.linear`` -> cuBLAS / cuBLASLt
  * attention           ``flash_attn`` 2.x (``flash_attn_qkvpacked_func``: causal, ``alibi_slopes``, fused dropout)
  * LayerNorm / GELU    ``F.layer_norm`` / ``F.gelu(approximate="tanh")`` (ATen kernels)
  * loss                ``F.cross_entropy`` on fp32 logits, 4096-row chunks (bounds the fp32 logits to 0.8 GB)
  * ZeRO-1              flat bf16 parameter / gradient buffers; ONE ``reduce_scatter_tensor`` over the flat gradient,
                        ``torch.optim.AdamW(fused=True)`` on the fp32 shard, ONE ``all_gather_into_tensor`` (NCCL)

None of this repo's kernels, ops or engine are on this path (``bench.py --impl baseline`` imports only this file and the
YAML model table).  It replaces round 1's baseline, which fed SDPA a dense additive mask, ran AdamW as ~10 eager
elementwise passes and issued one NCCL call per parameter region -- a straw man.

The reference itself (JAX/TPU, ``/root/reference/main_zero.py``) cannot run here; this is its algorithm on the stock CUDA
stack: all-reduce+slice -> reduce-scatter is the only liberty taken, in the baseline's favour.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F
import yaml


def _slopes(n: int):
    def pow2(m):
        base = 2.0 ** (-8.0 / m)
        return [base ** (i + 1) for i in range(m)]

    if n & (n - 1) == 0:
        return pow2(n)
    lo = 1 << (n.bit_length() - 1)
    return pow2(lo) + _slopes(2 * lo)[0::2][: n - lo]


class Block(nn.Module):
    def __init__(self, D, H, N, p):
        super().__init__()
        self.H, self.p = H, p
        self.ln1 = nn.Parameter(torch.ones(D))
        self.ln2 = nn.Parameter(torch.ones(D))
        self.qkv = nn.Parameter(torch.empty(3 * D, D).normal_(std=0.02))
        self.out = nn.Parameter(torch.empty(D, D).normal_(std=0.02 / math.sqrt(2 * N)))
        self.fc_in = nn.Parameter(torch.empty(4 * D, D).normal_(std=0.02))
        self.fc_out = nn.Parameter(torch.empty(D, 4 * D).normal_(std=0.02 / math.sqrt(2 * N)))
        self.register_buffer("slopes", torch.tensor(_slopes(H), dtype=torch.float32), persistent=False)

    def forward(self, x, train):
        from flash_attn import flash_attn_qkvpacked_func

        B, T, D = x.shape
        hd = D // self.H
        h = F.layer_norm(x, (D,), self.ln1, None, 1e-5)
        qkv = F.linear(h, self.qkv).view(B, T, 3, self.H, hd)
        a = flash_attn_qkvpacked_func(qkv, dropout_p=self.p if train else 0.0, softmax_scale=hd ** -0.5, causal=True,
                                      alibi_slopes=self.slopes).reshape(B, T, D)
        x = x + F.dropout(F.linear(a, self.out), self.p, train)
        h = F.layer_norm(x, (D,), self.ln2, None, 1e-5)
        h = F.linear(F.gelu(F.linear(h, self.fc_in), approximate="tanh"), self.fc_out)
        return x + F.dropout(h, self.p, train)


class GPT(nn.Module):
    def __init__(self, D, V, H, N, p):
        super().__init__()
        self.wte = nn.Parameter(torch.empty(V, D).normal_(std=0.02))
        self.blocks = nn.ModuleList([Block(D, H, N, p) for _ in range(N)])
        self.ln_f = nn.Parameter(torch.ones(D))

    def loss(self, tokens, train=True, chunk=4096):
        x = F.embedding(tokens.long(), self.wte)
        for b in self.blocks:
            x = b(x, train)
        D = x.shape[-1]
        h = F.layer_norm(x, (D,), self.ln_f, None, 1e-5)[:, :-1].reshape(-1, D)
        y = tokens[:, 1:].reshape(-1).long()
        total = h.new_zeros((), dtype=torch.float32)
        for s in range(0, h.shape[0], chunk):
            total = total + F.cross_entropy(F.linear(h[s:s + chunk], self.wte).float(), y[s:s + chunk], reduction="sum")
        return total / y.numel()


class BaselineTrainer:
    """``train_step(batch)`` = zero grads, ``accum`` micro-batches of forward+backward, reduce-scatter, clip + fused AdamW on
    the fp32 shard, bf16 cast, all-gather."""

    def __init__(self, model_name: str, model_cfg_path: str, device, rank: int, world: int, accum: int, seq: int,
                 lr=3e-4, wd=0.1, betas=(0.9, 0.95), eps=1e-8, clip=1.0, dropout=0.1):
        table = yaml.safe_load(open(model_cfg_path))
        e = table[model_name]
        torch.manual_seed(0)
        with torch.device(device):
            model = GPT(e["embedding_dim"], e["vocab_size"], e["num_head"], e["N"], float(e.get("dropout", dropout)))
        self.cfg = e
        self.model, self.device, self.rank, self.world, self.accum, self.seq, self.clip = model, device, rank, world, accum, seq, clip
        params = list(model.parameters())
        n = sum(p.numel() for p in params)
        pad = (-n) % (world * 64)
        self.flat_param = torch.zeros(n + pad, dtype=torch.bfloat16, device=device)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.shard = (n + pad) // world
        decay = torch.zeros(n + pad, dtype=torch.bool, device=device)
        off = 0
        master_full = torch.zeros(n + pad, dtype=torch.float32, device=device)
        for p in params:
            k = p.numel()
            master_full[off:off + k].copy_(p.detach().reshape(-1))
            decay[off:off + k] = p.ndim >= 2
            self.flat_param[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_param[off:off + k].view(p.shape)
            p.grad = self.flat_grad[off:off + k].view(p.shape)
            off += k
        s = rank * self.shard
        # two parameter groups (decay / no decay) as boolean-masked halves would need index ops; instead keep two master
        # tensors selected once: decayed elements and the rest of this rank's shard
        self.master = nn.Parameter(master_full[s:s + self.shard].clone())
        self.decay_mask = decay[s:s + self.shard].float()
        del master_full
        # weight decay is applied by hand (mask) so that one fused AdamW call covers the shard
        self.opt = torch.optim.AdamW([self.master], lr=lr, betas=betas, eps=eps, weight_decay=0.0, fused=True)
        self.lr, self.wd = lr, wd
        self.gshard = torch.empty(self.shard, dtype=torch.bfloat16, device=device)

    def num_params(self):
        return sum(p.numel() for p in self.model.parameters())

    def train_step(self, batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.model.train()
        self.flat_grad.zero_()
        rows = batch.reshape(-1, self.seq)
        mbs = rows.view(self.accum, rows.shape[0] // self.accum, self.seq)
        loss_sum = None
        for i in range(self.accum):
            loss = self.model.loss(mbs[i], True)
            loss.backward()
            loss_sum = loss.detach() if loss_sum is None else loss_sum + loss.detach()
        if self.world > 1:
            dist.reduce_scatter_tensor(self.gshard, self.flat_grad, op=dist.ReduceOp.SUM)
            g = self.gshard
        else:
            g = self.flat_grad[: self.shard]
        grad = g.float().mul_(1.0 / (self.world * self.accum)).clamp_(-self.clip, self.clip)
        with torch.no_grad():
            self.master.mul_(1.0 - self.lr * self.wd * self.decay_mask)   # decoupled weight decay on ndim >= 2 parameters
        self.master.grad = grad
        self.opt.step()
        s = self.rank * self.shard
        with torch.no_grad():
            self.flat_param[s:s + self.shard].copy_(self.master)
            if self.world > 1:
                dist.all_gather_into_tensor(self.flat_param, self.flat_param[s:s + self.shard])
        return {"train/loss": loss_sum / self.accum, "lr": self.lr}
