"""Factory tests (model of reference tests/test_model_factory.py:19-32)."""
import pytest
import torch

from zero_transformer_b200.config import DEFAULT_MODEL_CFG, load_yaml
from zero_transformer_b200.models.gpt import GPTConfig, model_getter


def test_factory_honours_yaml():
    m = model_getter("test")
    assert m.N == load_yaml(DEFAULT_MODEL_CFG)["test"]["N"] == len(m.blocks)


def test_unknown_name_raises():
    with pytest.raises(AssertionError):
        model_getter("does_not_exist")


def test_invalid_dtype_raises():
    with pytest.raises(AssertionError):
        model_getter("test", dtype=torch.float64)


def test_return_cfg_has_reference_keys():
    _, c = model_getter("test", return_cfg=True)
    for k in ("embedding_dim", "vocab_size", "num_head", "block_size", "dropout", "N", "alibi_attn"):
        assert k in c


@pytest.mark.parametrize("name,count", [("417m", 417_043_968), ("760m", 756_819_456), ("1_3b", 1_311_082_496)])
def test_named_model_sizes(name, count):
    e = load_yaml(DEFAULT_MODEL_CFG)[name]
    c = GPTConfig(**{k: v for k, v in dict(e).items() if k in GPTConfig.__dataclass_fields__})
    assert c.num_params() == count
