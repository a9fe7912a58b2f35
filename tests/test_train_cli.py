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


def _uneven_worker(rank, world, port, root, out):
    import os

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from zero_transformer_b200.config import DEFAULT_MODEL_CFG, load_config
    from zero_transformer_b200.parallel import dist as pdist
    from zero_transformer_b200.train import Trainer

    env = pdist.setup_dp(device="cpu")
    cfg = load_config(None, ["model.size=test", "training.batch_size=4", "training.train_context=16", "data.max_context=32",
                             "training.gradient_accumulation_steps=1", "training.evaluation_frequency=2", "training.maximum_evaluation_steps=5",
                             "training.warmup_steps=1", "training.total_steps=50", "training.max_epochs=1", "data.corpus=files",
                             f"data.bucket_path={root}", f"data.index_path_train={root}/train.index",
                             f"data.index_path_validation={root}/val.index", "data.shuffle_buffer=1"])
    tr = Trainer(cfg, DEFAULT_MODEL_CFG, env, log_metrics=False, synthetic=False)
    last = tr.fit()
    if rank == 0:
        import json

        json.dump({"final_step": tr.final_step, "keys": sorted(last)}, open(out, "w"))
    pdist.shutdown()


def test_ranks_stop_together_when_one_runs_out_of_data(tmp_path):
    """Real shards give ranks different batch counts (striding + drop_last): rank 0 gets 3 train batches, rank 1 only 2; both
    must leave the loop after 2 steps (and the validation loop after its common prefix) instead of hanging in a collective."""
    import json
    import socket

    import numpy as np
    import torch.multiprocessing as mp

    rng = np.random.default_rng(0)
    np.save(tmp_path / "a.npy", rng.integers(0, 256, size=(6, 32), dtype=np.int32))   # rank 0: 6 rows -> 3 batches of 2
    np.save(tmp_path / "b.npy", rng.integers(0, 256, size=(5, 32), dtype=np.int32))   # rank 1: 5 rows -> 2 batches
    (tmp_path / "train.index").write_text("a.npy\nb.npy\n")
    np.save(tmp_path / "va.npy", rng.integers(0, 256, size=(2, 32), dtype=np.int32))  # rank 0: 2 validation batches of 1
    np.save(tmp_path / "vb.npy", rng.integers(0, 256, size=(1, 32), dtype=np.int32))  # rank 1: 1
    (tmp_path / "val.index").write_text("va.npy\nvb.npy\n")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "res.json")
    mp.spawn(_uneven_worker, args=(2, port, str(tmp_path), out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res["final_step"] == 2 and "validation/loss" in res["keys"]
