"""Sampling utilities + KV-cache parity (model of torch_compatability/test_torch_models.py:42-91)."""
import torch

from zero_transformer_b200.generate import (ByteTokenizer, generate_from_prompt, generate_text, process_logits, top_k_logits,
                                            top_p_logits)
from zero_transformer_b200.models.gpt import model_getter


def test_top_k_and_top_p():
    lg = torch.tensor([[1.0, 3.0, 2.0, 0.0]])
    k = top_k_logits(lg.clone(), 2)
    assert torch.isinf(k[0, 0]) and torch.isinf(k[0, 3]) and k[0, 1] == 3 and k[0, 2] == 2
    p = top_p_logits(torch.log(torch.tensor([[0.5, 0.3, 0.15, 0.05]])), 0.7)
    assert (~torch.isinf(p)).sum() == 2
    assert torch.equal(top_p_logits(lg, 0.0), lg) and torch.equal(top_k_logits(lg, 0), lg)


def test_repetition_penalty_and_temperature():
    lg = torch.tensor([[2.0, -2.0, 1.0]])
    out = process_logits(lg.clone(), torch.tensor([[0, 1]]), temperature=0.5, repetition_penalty=2.0)
    torch.testing.assert_close(out, torch.tensor([[2.0, -8.0, 2.0]]))


def test_kv_cache_grows_and_matches_full_forward():
    m = model_getter("test").eval()
    x = torch.randint(0, 256, (1, 12))
    _, past = m(x[:, :8], use_cache=True)
    assert past[0][0].shape == (1, 4, 8, 16)
    for t in range(8, 12):
        lg, past = m(x[:, t:t + 1], use_cache=True, past_states=past)
        assert past[0][0].shape[2] == t + 1
        torch.testing.assert_close(lg[:, -1], m(x[:, :t + 1])[:, -1], rtol=1e-4, atol=1e-5)


def test_forward_longer_than_block_size():
    # ALiBi extrapolation: T=64 > block_size=32 (reference test_torch_models.py:198-203)
    m = model_getter("test").eval()
    assert m(torch.randint(0, 256, (1, 64))).shape == (1, 64, 256)


def test_greedy_generation_is_deterministic_and_stops_on_eos():
    m = model_getter("test").eval()
    a = list(generate_from_prompt(m, [1, 2, 3], 10, sample=False))
    b = list(generate_from_prompt(m, [1, 2, 3], 10, sample=False))
    assert a == b and len(a) == 10
    c = list(generate_from_prompt(m, [1, 2, 3], 10, sample=False, eos_token_id=a[2]))
    assert c == a[: a.index(a[2]) + 1]
    assert isinstance(generate_text(m, ByteTokenizer(), "ab", 4, seed=0), str)
    long = list(generate_from_prompt(m, list(range(40)), 5, sample=False))  # prompt longer than the window
    assert len(long) == 5


def test_model_generate_matches_cached_greedy():
    m = model_getter("test").eval()
    ctx = [3, 1, 4, 1, 5]
    full = m.generate(ctx, max_length=12)
    assert full.shape == (1, 12) and full[0, :5].tolist() == ctx
    cached = list(generate_from_prompt(m, ctx, 7, sample=False))
    assert full[0, 5:].tolist() == cached
