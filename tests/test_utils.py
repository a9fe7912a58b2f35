"""Loss / schedule / config / misc utilities (model of reference tests/test_utils.py:11-57)."""
import math

import pytest
import torch

from zero_transformer_b200.config import Config, flatten_dict, load_config
from zero_transformer_b200.ops import reference as R
from zero_transformer_b200.optim.schedule import WarmupCosine
from zero_transformer_b200.utils.extend_params import create_mapping, extend_params
from zero_transformer_b200.utils.metrics import compute_tokens_seen


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_loss_is_fp32(dtype):
    logits = torch.randn(4, 7, 11).to(dtype)
    labels = torch.randint(0, 11, (4, 7))
    assert R.cross_entropy(logits, labels).dtype == torch.float32


def test_loss_golden_uniform():
    # uniform logits -> loss = log(V) exactly
    logits = torch.zeros(3, 5, 50304)
    labels = torch.randint(0, 50304, (3, 5))
    assert R.cross_entropy(logits, labels).item() == pytest.approx(math.log(50304), rel=1e-6)


def test_loss_matches_onehot_formula():
    # the reference computes -mean(sum(onehot * log_softmax)) (losses.py:21-23)
    logits = torch.randn(2, 6, 13)
    labels = torch.randint(0, 13, (2, 6))
    onehot = torch.nn.functional.one_hot(labels, 13).float()
    ref = -(onehot * torch.log_softmax(logits.float(), -1)).sum(-1).mean()
    torch.testing.assert_close(R.cross_entropy(logits, labels), ref)


def test_warmup_cosine_schedule():
    s = WarmupCosine(peak_value=3e-4, warmup_steps=2000, decay_steps=143000, end_value=3e-5)
    assert s(0) == 0.0
    assert s(1000) == pytest.approx(1.5e-4)
    assert s(2000) == pytest.approx(3e-4)
    mid = 2000 + (143000 - 2000) // 2
    assert s(mid) == pytest.approx((3e-4 + 3e-5) / 2, rel=1e-3)
    assert s(143000) == pytest.approx(3e-5)
    assert s(10 ** 7) == pytest.approx(3e-5)
    assert all(s(i) >= s(i + 1) for i in range(2000, 143000, 997))


def test_flatten_dict():
    d = {"a": {"b": 1, "c": {"d": 2}}, "e": 3}
    assert flatten_dict(d) == {"a.b": 1, "a.c.d": 2, "e": 3}
    assert flatten_dict(Config(d)) == {"a.b": 1, "a.c.d": 2, "e": 3}


def test_config_overrides_and_validation(tmp_path):
    cfg = load_config(None, ["training.batch_size=32", "model.size=test"])
    assert cfg.training.batch_size == 32 and cfg.model.size == "test"
    assert cfg.training.peak_learning_rate == pytest.approx(3e-4)
    with pytest.raises(ValueError):
        load_config(None, ["training.train_context=1000"])  # 2048 % 1000 != 0
    with pytest.raises(AttributeError):
        _ = cfg.training.nope


def test_tokens_seen_is_exact():
    assert compute_tokens_seen(10, 256, 2048) == 10 * 256 * 2048


def test_extend_params_doubles_depth():
    src = {"wte.weight": torch.randn(4, 2), "ln_f.weight": torch.ones(2)}
    for i in range(3):
        src[f"blocks.{i}.attn.qkv.weight"] = torch.full((2, 2), float(i))
    out = extend_params(src, 6)
    assert create_mapping(3, 6) == {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 2}
    for j in range(6):
        assert torch.all(out[f"blocks.{j}.attn.qkv.weight"] == j // 2)
    assert torch.equal(out["wte.weight"], src["wte.weight"])
    with pytest.raises(ValueError):
        extend_params(src, 7)
