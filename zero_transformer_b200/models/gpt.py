"""GPT-2-style ALiBi decoder.

Architecture parity with the reference flax model (``src/models/GPT.py:16-113``,
``src/models/layers.py:47-191``): token embedding only (no learned positions),
pre-LayerNorm blocks, bias-free projections and LayerNorms, tanh-GELU 4x MLP,
dropout on attention probabilities / attention output / MLP output, final
LayerNorm, LM head tied to the embedding, shifted next-token cross entropy.

B200-first differences in *layout*, not in math:
  * q/k/v are ONE ``(3D, D)`` weight so the projection is a single tcgen05 GEMM
    (and a single all-gather-fused GEMM in the ZeRO-1 path); the exporter
    splits it back into ``key/query/value`` for the reference ``.pth`` contract.
  * weights are stored ``(out, in)`` = K-major for both GEMM operands.
  * every op goes through ``zero_transformer_b200.ops`` which runs hand-written
    sm_100a kernels on CUDA tensors and PyTorch reference code on CPU.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from .. import ops
from ..config import DEFAULT_MODEL_CFG, load_yaml


@dataclass
class GPTConfig:
    embedding_dim: int
    vocab_size: int
    num_head: int
    block_size: int
    N: int
    dropout: float = 0.0
    alibi_attn: bool = True
    # --- additions (SURVEY 7.1 decisions) ---
    attn_scale: str = "head_dim"  # "head_dim" | "seq_len" (reference flax quirk)
    ln_eps: float = 1e-5          # torch default; flax default is 1e-6
    gelu: str = "tanh"            # flax nn.gelu default; torch consumer uses erf

    @property
    def head_dim(self) -> int:
        return self.embedding_dim // self.num_head

    def num_params(self) -> int:
        D, V, N = self.embedding_dim, self.vocab_size, self.N
        return V * D + N * (12 * D * D + 2 * D) + D

    def flops_per_token(self, seq_len: int, causal: bool = True) -> float:
        """Training FLOPs/token: 6*P (fwd+bwd GEMMs incl. tied LM head) + attention (QK^T and PV, fwd+bwd).
        ``causal=True`` (default, used for every MFU this repo reports) counts only the lower-triangular half of the
        score matrix that a causal kernel has to compute: 6*N*T*D; ``causal=False`` gives the dense 12*N*T*D that is
        often quoted and overstates the work by ~3.5 % at T=1024 on the 1.3B model."""
        attn = (6.0 if causal else 12.0) * self.N * seq_len * self.embedding_dim
        return 6.0 * self.num_params() + attn


class LayerNorm(nn.Module):
    """Bias-free LayerNorm (reference ``GPT.py:42,49,98``)."""

    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        return ops.layernorm(x, self.weight, self.eps)

    def forward_res(self, x):
        """-> (normalised x, x for the residual branch)"""
        return ops.layernorm_res(x, self.weight, self.eps)


class Linear(nn.Module):
    """Bias-free ``(out, in)`` weight holder; the matmul happens in ``ops``."""

    def __init__(self, d_in: int, d_out: int, std: float):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(d_out, d_in))
        self.std = std
        nn.init.normal_(self.weight, mean=0.0, std=std)

    def forward(self, x):
        return ops.linear(x, self.weight)


class MLPBlock(nn.Module):
    """fc_in -> GELU -> fc_out -> dropout (reference ``layers.py:47-77``).
    ``forward`` returns ``residual + dropout(mlp(x))`` so the last GEMM can
    carry the dropout+residual epilogue."""

    def __init__(self, cfg: GPTConfig):
        super().__init__()
        D = cfg.embedding_dim
        self.fc_in = Linear(D, 4 * D, 0.02)
        self.fc_out = Linear(4 * D, D, 0.02 / math.sqrt(2 * cfg.N))
        self.p = cfg.dropout
        self.gelu = cfg.gelu

    def forward(self, x, residual, train: bool):
        return ops.mlp(x, self.fc_in.weight, self.fc_out.weight, residual, self.p, train, self.gelu)


class CausalAttention(nn.Module):
    """Multi-head causal self-attention with optional ALiBi
    (reference ``layers.py:80-191``)."""

    def __init__(self, cfg: GPTConfig):
        super().__init__()
        D = cfg.embedding_dim
        assert D % cfg.num_head == 0, "embedding_dim must be divisible by num_head"
        self.cfg = cfg
        self.qkv = Linear(D, 3 * D, 0.02)  # rows: [Wq ; Wk ; Wv]
        self.out = Linear(D, D, 0.02 / math.sqrt(2 * cfg.N))
        if cfg.alibi_attn:
            self.register_buffer("slopes", torch.tensor(ops.get_slopes(cfg.num_head), dtype=torch.float32), persistent=False)
        else:
            self.slopes = None

    def _apply(self, fn, recurse=True):
        # ``model.to(bfloat16)`` must not round the ALiBi slopes (2^-8/H ... are not all bf16-exact for H not a power of
        # two): the buffer follows the module's device but always stays fp32.
        super()._apply(fn, recurse)
        if self.slopes is not None and self.slopes.dtype != torch.float32:
            self.slopes = torch.tensor(ops.get_slopes(self.cfg.num_head), dtype=torch.float32, device=self.slopes.device)
        return self

    def forward(self, x, residual, train: bool, layer_past=None, use_cache: bool = False):
        cfg = self.cfg
        T = x.shape[1]
        qkv = ops.linear(x, self.qkv.weight)
        scale = ops.default_attn_scale(cfg.head_dim, T, cfg.attn_scale)
        present = None
        if use_cache:
            a, present = ops.attention(qkv, self.slopes, cfg.num_head, scale, 0.0, False,
                                       kv_cache=layer_past if layer_past is not None else (None, None))
        else:
            a = ops.attention(qkv, self.slopes, cfg.num_head, scale, cfg.dropout, train)
        y = ops.linear_dropout_add(a, self.out.weight, residual, cfg.dropout, train)
        return y, present


class TransformerBlock(nn.Module):
    """Pre-LN block: ``x += attn(ln1(x)); x += mlp(ln2(x))`` (reference ``GPT.py:16-50``)."""

    def __init__(self, cfg: GPTConfig):
        super().__init__()
        self.ln1 = LayerNorm(cfg.embedding_dim, cfg.ln_eps)
        self.attn = CausalAttention(cfg)
        self.ln2 = LayerNorm(cfg.embedding_dim, cfg.ln_eps)
        self.mlp = MLPBlock(cfg)

    def forward(self, x, train: bool = False, layer_past=None, use_cache: bool = False):
        h, xr = self.ln1.forward_res(x)
        x, present = self.attn(h, xr, train, layer_past, use_cache)
        h, xr = self.ln2.forward_res(x)
        x = self.mlp(h, xr, train)
        return x, present


class Transformer(nn.Module):
    """Full model.  ``model(x)`` -> logits ``(B, T, V)``; ``model(x, labels)`` ->
    ``(logits, loss)`` like reference ``GPT.py:68-113``; ``model.loss(x)`` is the
    training entry: LM-head GEMM + cross entropy as one autograd node that keeps a single bf16
    logits buffer (no fp32 logits, no softmax tensor) and turns it into the gradient in place."""

    def __init__(self, cfg: GPTConfig):
        super().__init__()
        self.cfg = cfg
        self.wte = nn.Embedding(cfg.vocab_size, cfg.embedding_dim)
        nn.init.normal_(self.wte.weight, mean=0.0, std=0.02)
        self.blocks = nn.ModuleList([TransformerBlock(cfg) for _ in range(cfg.N)])
        self.ln_f = LayerNorm(cfg.embedding_dim, cfg.ln_eps)
        # optional ZeRO-1 engine hooks: called with the block index right before block i runs in
        # forward (parameter prefetch wait) and right after its backward finished (bucket ready)
        self.block_forward_pre_hook = None
        self.block_backward_done_hook = None

    # -- the reference keeps these as constructor attributes
    @property
    def N(self) -> int:
        return self.cfg.N

    @property
    def block_size(self) -> int:
        return self.cfg.block_size

    def hidden(self, x: torch.Tensor, train: bool = False, past_states=None, use_cache: bool = False):
        if train:
            ops.begin_forward(self.wte.weight.device)
        h = ops.embedding(x, self.wte.weight)
        presents = []
        for i, blk in enumerate(self.blocks):
            past = past_states[i] if past_states is not None else None
            if self.block_forward_pre_hook is not None:
                self.block_forward_pre_hook(i)
            if self.block_backward_done_hook is not None and h.requires_grad:
                h.register_hook(lambda g, i=i: self.block_backward_done_hook(i))
            h, present = blk(h, train, past, use_cache)
            presents.append(present)
        return self.ln_f(h), presents

    def loss(self, x: torch.Tensor, labels: Optional[torch.Tensor] = None, train: Optional[bool] = None) -> torch.Tensor:
        train = self.training if train is None else train
        h, _ = self.hidden(x, train)
        return ops.lm_head_loss(h, self.wte.weight, x if labels is None else labels)

    def forward(
        self,
        x: torch.Tensor,
        labels: Optional[torch.Tensor] = None,
        train: Optional[bool] = None,
        use_cache: bool = False,
        past_states=None,
    ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, list]]:
        train = self.training if train is None else train
        h, presents = self.hidden(x, train, past_states, use_cache)
        logits = ops.linear(h, self.wte.weight)
        if labels is not None:
            loss = ops.cross_entropy(logits[:, :-1], labels[:, 1:])
            return logits, loss
        if use_cache:
            return logits, presents
        return logits

    @torch.no_grad()
    def generate(self, context, max_length: int, sample: bool = False) -> torch.Tensor:
        """Greedy / sampled continuation without a KV cache -- the lm-eval-harness hook of the
        reference ``GPT2.generate`` (``torch_compatability/GPT2.py:354-400``).  ``max_length`` counts
        context + generated tokens; the context window is cropped to ``block_size``."""
        dev = self.wte.weight.device
        x = torch.as_tensor(context, dtype=torch.long, device=dev).view(1, -1)
        for _ in range(max_length - x.shape[1]):
            x_cond = x[:, -self.cfg.block_size:]
            probs = torch.softmax(self.forward(x_cond, train=False)[:, -1, :].float(), dim=-1)
            nxt = torch.multinomial(probs, 1) if sample else probs.argmax(-1, keepdim=True)
            x = torch.cat([x, nxt], dim=1)
        return x

    def decay_mask(self) -> dict:
        """``True`` for parameters that receive weight decay: everything with
        ndim >= 2 (reference mask ``main_zero.py:155-158``)."""
        return {n: p.ndim >= 2 for n, p in self.named_parameters()}


def _cfg_from_entry(entry, dtype) -> GPTConfig:
    allowed = set(GPTConfig.__dataclass_fields__)
    kw = {k: v for k, v in dict(entry).items() if k in allowed}
    if "block_size" not in kw and "num_ctx" in entry:  # torch-side table spelling
        kw["block_size"] = entry["num_ctx"]
    return GPTConfig(**kw)


def model_getter(
    model_size: str,
    config_path: str = DEFAULT_MODEL_CFG,
    return_cfg: bool = False,
    dtype: torch.dtype = torch.float32,
    device: Union[str, torch.device] = "cpu",
    **overrides,
):
    """Build a named model from the YAML table (reference ``GPT.py:116-137``;
    same asserts as ``tests/test_model_factory.py:19-32``)."""
    table = load_yaml(config_path)
    assert model_size in list(table.keys()), "Invalid model name provided"
    assert dtype in (torch.float16, torch.bfloat16, torch.float32), "Invalid dtype provided"
    cfg = _cfg_from_entry(table[model_size], dtype)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    with torch.device(device):  # init directly on the target device (a 1.3B CPU init takes ~1 min)
        model = Transformer(cfg)
    model = model.to(dtype=dtype)
    if return_cfg:
        return model, asdict(cfg)
    return model
