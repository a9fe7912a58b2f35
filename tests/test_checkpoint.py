"""Checkpoint / resume / warm-start (untested in the reference -- SURVEY 4)."""
import os

import torch

from zero_transformer_b200.models.gpt import GPTConfig, Transformer
from zero_transformer_b200.optim.schedule import WarmupCosine
from zero_transformer_b200.parallel.zero1 import Zero1Engine
from zero_transformer_b200.utils import checkpoint as ckpt

LR = WarmupCosine(peak_value=1e-2, warmup_steps=2, decay_steps=50, end_value=1e-3)


def _mk(N=2, seed=0):
    torch.manual_seed(seed)
    m = Transformer(GPTConfig(embedding_dim=32, vocab_size=64, num_head=2, block_size=16, N=N, dropout=0.0))
    return m, Zero1Engine(m, lr_fn=LR, backend="nccl")


def _steps(m, e, n, seed=5):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n):
        e.zero_grad()
        m.loss(torch.randint(0, 64, (4, 16), generator=g), train=False).backward()
        e.step()


def test_save_resume_roundtrip_is_bit_exact(tmp_path):
    m, e = _mk()
    _steps(m, e, 3)
    ckpt.save_checkpoint(e, str(tmp_path), 3)
    assert os.path.exists(tmp_path / "params" / "params_3.pt") and os.path.exists(tmp_path / "optimizer" / "optimizer_3.pt")
    _steps(m, e, 2, seed=9)
    m2, e2 = _mk(seed=123)
    assert ckpt.resume(e2, str(tmp_path)) == 3 and e2.step_count == 3
    _steps(m2, e2, 2, seed=9)
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), n
    assert torch.equal(e.exp_avg_sq, e2.exp_avg_sq)


def test_sharded_format_roundtrip(tmp_path):
    m, e = _mk()
    _steps(m, e, 2)
    ckpt.save_checkpoint(e, str(tmp_path), 2, sharded=True)
    m2, e2 = _mk(seed=7)
    assert ckpt.resume(e2, str(tmp_path)) == 2
    assert torch.equal(e.master, e2.master) and torch.equal(e.exp_avg, e2.exp_avg)


def test_keep_last_five(tmp_path):
    m, e = _mk()
    for s in range(1, 8):
        ckpt.save_checkpoint(e, str(tmp_path), s, keep=5)
    assert sorted(os.listdir(tmp_path / "params")) == [f"params_{s}.pt" for s in range(3, 8)]
    assert ckpt.latest_step(str(tmp_path)) == 7
    ckpt.wipe_checkpoints(str(tmp_path))
    assert ckpt.latest_step(str(tmp_path)) is None


def test_warm_start_with_depth_extension(tmp_path):
    m, e = _mk(N=2)
    _steps(m, e, 2)
    ckpt.save_checkpoint(e, str(tmp_path), 2)
    m4, e4 = _mk(N=4, seed=3)
    ckpt.warm_start(e4, str(tmp_path), extend_depth=True)
    # the optimizer counter (Adam bias correction + LR schedule) travels with the moments, as in the reference
    # (ScaleByAdamState.count / ScaleByScheduleState.count, main_zero.py:105-139); the *training* step stays with the Trainer
    assert e4.step_count == 2
    src = dict(m.named_parameters())
    dst = dict(m4.named_parameters())
    for j in range(4):
        assert torch.equal(dst[f"blocks.{j}.mlp.fc_in.weight"], src[f"blocks.{j // 2}.mlp.fc_in.weight"])
    assert torch.equal(dst["wte.weight"], src["wte.weight"])
