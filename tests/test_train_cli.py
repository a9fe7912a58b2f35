"""End-to-end driver on CPU: train -> checkpoint -> resume -> export -> generate."""
import json

import torch

from zero_transformer_b200 import generate, train
from zero_transformer_b200.utils import convert


def _args(tmp_path, extra=()):
    return ["--synthetic", "--device", "cpu", "model.size=test", "training.batch_size=4", "training.train_context=16",
            "data.max_context=32", "training.gradient_accumulation_steps=2", "training.evaluation_frequency=3",
            "training.maximum_evaluation_steps=2", "training.warmup_steps=2", "training.total_steps=6",
            f"data.bucket_path={tmp_path}", *extra]


def test_train_resume_export_generate(tmp_path):
    assert train.main(_args(tmp_path, ["--max-steps", "4"])) == 0
    ck = tmp_path / "checkpoints"
    assert (ck / "params" / "params_3.pt").exists() and (ck / "optimizer" / "optimizer_3.pt").exists()
    recs = [json.loads(l) for l in open(ck / "metrics.jsonl")]
    steps = [r for r in recs if "step" in r]
    assert [r["step"] for r in steps] == [1, 2, 3, 4]
    for k in ("train/loss", "train/ppl", "Learning Rate", "Tokens Seen (B)", "Train Sequence Length"):
        assert k in steps[0]
    assert "validation/loss" in steps[2] and steps[0]["Learning Rate"] == 0.0
    assert steps[3]["train/loss"] < steps[0]["train/loss"] + 0.5

    assert train.main(["--resume", *_args(tmp_path)]) == 0  # resumes at 3, runs to total_steps=6
    recs = [json.loads(l) for l in open(ck / "metrics.jsonl")]
    assert [r["step"] for r in recs if "step" in r][-3:] == [4, 5, 6]
    assert (ck / "params" / "params_6.pt").exists()

    out = tmp_path / "m.pth"
    assert convert.main(["extract", "--ckpt-dir", str(ck), "--out", str(tmp_path / "p.pt")]) == 0
    assert convert.main(["to-torch", "--model-name", "test", "--params", str(tmp_path / "p.pt"), "--torch-path", str(out)]) == 0
    sd = torch.load(out, weights_only=False)
    assert len(sd) == 18 * 2 + 4
    model = generate.model_creator("test", str(out), device="cpu")
    assert len(list(generate.generate_from_prompt(model, [5, 6], 4, sample=False))) == 4
