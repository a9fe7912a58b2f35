"""Layer-level tests (model of reference tests/test_model_components.py:12-262:
construct + forward shape checks, train-vs-eval dropout, internal loss == external CE),
plus numeric checks the reference lacks (causality, ALiBi last-row equivalence)."""
import math

import pytest
import torch

from zero_transformer_b200 import ops
from zero_transformer_b200.models.gpt import (CausalAttention, GPTConfig, MLPBlock, Transformer, TransformerBlock)
from zero_transformer_b200.ops import reference as R


def cfg(**kw):
    base = dict(embedding_dim=128, vocab_size=256, num_head=8, block_size=512, N=2, dropout=0.1, alibi_attn=True)
    base.update(kw)
    return GPTConfig(**base)


def test_mlp_shape_and_residual():
    c = cfg()
    mlp = MLPBlock(c).eval()
    x = torch.randn(3, 17, 128)
    y = mlp(x, torch.zeros_like(x), False)
    assert y.shape == x.shape
    y2 = mlp(x, x, False)
    torch.testing.assert_close(y2, y + x)


@pytest.mark.parametrize("alibi", [True, False])
def test_attention_shape(alibi):
    c = cfg(alibi_attn=alibi)
    attn = CausalAttention(c).eval()
    x = torch.randn(2, 64, 128)
    y, present = attn(x, torch.zeros_like(x), False)
    assert y.shape == x.shape and present is None


def test_attention_dropout_train_vs_eval():
    c = cfg(dropout=0.5)
    attn = CausalAttention(c)
    x = torch.randn(2, 16, 128)
    e1, _ = attn(x, x, False)
    e2, _ = attn(x, x, False)
    t1, _ = attn(x, x, True)
    torch.testing.assert_close(e1, e2)
    assert not torch.allclose(e1, t1)


def test_attention_is_causal():
    c = cfg(dropout=0.0)
    attn = CausalAttention(c).eval()
    x = torch.randn(1, 32, 128)
    y1, _ = attn(x, torch.zeros_like(x), False)
    x2 = x.clone()
    x2[:, 20:] = torch.randn(1, 12, 128)
    y2, _ = attn(x2, torch.zeros_like(x2), False)
    torch.testing.assert_close(y1[:, :20], y2[:, :20])


def test_alibi_last_row_form_is_equivalent():
    # reference create_mask (layers.py:33-44) keeps only the last row of the distance matrix;
    # after softmax it equals the canonical slope*(j-i) bias.
    H, T = 4, 16
    slopes = torch.tensor(R.get_slopes(H))
    full = R.alibi_bias(slopes, T, T)                      # (H,T,T)
    last = full[:, T - 1: T, :].expand(H, T, T)            # last-row form broadcast to every query
    s = torch.randn(H, T, T)
    mask = torch.tril(torch.ones(T, T, dtype=torch.bool))
    a = torch.softmax((s + full).masked_fill(~mask, float("-inf")), -1)
    b = torch.softmax((s + last).masked_fill(~mask, float("-inf")), -1)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_slopes_match_published_values():
    assert R.get_slopes(8) == pytest.approx([2 ** -(i + 1) for i in range(8)])
    s12 = R.get_slopes(12)
    assert len(s12) == 12 and s12[:8] == pytest.approx([2 ** -(i + 1) for i in range(8)])
    assert s12[8:] == pytest.approx([2 ** -(0.5 + i) for i in range(4)])


def test_block_shape():
    blk = TransformerBlock(cfg()).eval()
    x = torch.randn(2, 32, 128)
    y, _ = blk(x, False)
    assert y.shape == x.shape


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_transformer_logits_shape(dtype):
    m = Transformer(cfg()).to(dtype).eval()
    x = torch.randint(0, 256, (2, 48))
    out = m(x)
    assert out.shape == (2, 48, 256) and out.dtype == dtype


def test_internal_loss_matches_external_ce():
    m = Transformer(cfg(dropout=0.0)).eval()
    x = torch.randint(0, 256, (2, 40))
    logits, loss = m(x, labels=x)
    ext = R.cross_entropy(logits[:, :-1], x[:, 1:])
    torch.testing.assert_close(loss, ext)
    torch.testing.assert_close(m.loss(x), ext, rtol=1e-5, atol=1e-5)
    assert loss.dtype == torch.float32


def test_attn_scale_switch():
    c1, c2 = cfg(dropout=0.0), cfg(dropout=0.0, attn_scale="seq_len")
    assert ops.default_attn_scale(16, 64, c1.attn_scale) == pytest.approx(0.25)
    assert ops.default_attn_scale(16, 64, c2.attn_scale) == pytest.approx(0.125)


def test_init_statistics():
    c = cfg(embedding_dim=256, N=4)
    m = Transformer(c)
    assert m.blocks[0].attn.qkv.weight.std().item() == pytest.approx(0.02, rel=0.1)
    assert m.blocks[0].mlp.fc_out.weight.std().item() == pytest.approx(0.02 / math.sqrt(8), rel=0.1)
    assert torch.all(m.ln_f.weight == 1)
    assert sum(p.numel() for p in m.parameters()) == c.num_params()
