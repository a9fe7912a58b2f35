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


def test_make_tables_pairs_arms_per_call_and_fills_baseline_md(tmp_path, monkeypatch):
    """bench/make_tables.py: both arms of a ratio come from the same sweep file (= same box), the latest call wins, and
    BASELINE.md's generated block lists every GPU count that has records."""
    import importlib
    import json
    import sys

    def rec(impl, n, value, model="1_3b", rows=32, accum=2, plane=None):
        return {"impl": impl, "value": value, "ms_per_step": 1.0, "n_gpus": n, "e2e": {"value": value},
                "config": {"model": f"gpt2-{model} ALiBi", "global_batch": rows * n, "seq_len": 1024, "grad_accum": accum,
                           "parallelism": f"dp{n}+zero1", "data_plane": plane or "nvls"}, "clocks": {"sm_mhz": 1500}}

    prof = tmp_path / "profiles"
    prof.mkdir()
    (prof / "sweep_n1_call1.jsonl").write_text("\n".join(json.dumps(r) for r in [rec("ours", 1, 100.0), rec("baseline", 1, 80.0)]))
    (prof / "sweep_n1_call2.jsonl").write_text("\n".join(json.dumps(r) for r in [rec("ours", 1, 130.0), rec("baseline", 1, 100.0), "garbage line"]))
    (prof / "sweep_n8_call3.jsonl").write_text("\n".join(json.dumps(r) for r in [rec("ours", 8, 1000.0), rec("baseline", 8, 800.0), rec("ours", 8, 990.0, plane="p2p")]))
    md = tmp_path / "BASELINE.md"
    md.write_text("head\n<!-- results:begin -->\nold\n<!-- results:end -->\ntail\n")
    monkeypatch.setenv("ZTB_PROFILES_DIR", str(prof))
    monkeypatch.setenv("ZTB_BASELINE_MD", str(md))
    sys.modules.pop("bench.make_tables", None)
    mt = importlib.import_module("bench.make_tables")
    mt.main()
    base = json.load(open(prof / "baseline_arm.json"))
    assert base["1_3b/seq1024/rows32/accum2/n1"]["value"] == 100.0 and base["1_3b/seq1024/rows32/accum2/n8"]["value"] == 800.0
    res = (prof / "RESULTS.md").read_text()
    assert "1.30x" in res and "1.25x" in res and "/p2p" in res and "| 8 | 1.0k | 0.962 |" in res
    out = md.read_text()
    assert "old" not in out and "1 / 8" in out and "1.30x / 1.25x" in out and out.startswith("head") and out.rstrip().endswith("tail")
    sys.modules.pop("bench.make_tables", None)


def test_attention_dq_mode_auto_rule(monkeypatch):
    """bf16 dQ accumulation: auto = up to 2048 tokens and not in deterministic mode; "0" / "1" force."""
    from zero_transformer_b200.ops import attention_cuda as A
    from zero_transformer_b200.ops import cuda as K

    monkeypatch.setattr(K.CONFIG, "attn_dq_bf16", "auto", raising=False)
    monkeypatch.setattr(K.CONFIG, "deterministic", False, raising=False)
    assert A._dq_bf16(1024) and A._dq_bf16(2048) and not A._dq_bf16(4096)
    monkeypatch.setattr(K.CONFIG, "deterministic", True, raising=False)
    assert not A._dq_bf16(1024)
    monkeypatch.setattr(K.CONFIG, "attn_dq_bf16", "1", raising=False)
    assert A._dq_bf16(8192)
    monkeypatch.setattr(K.CONFIG, "attn_dq_bf16", "0", raising=False)
    assert not A._dq_bf16(128)
    monkeypatch.setattr(K.CONFIG, "attn_dq_bf16", True, raising=False)   # benches set a bool
    assert A._dq_bf16(8192)
