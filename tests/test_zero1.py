"""ZeRO-1 engine: layout, single-rank equivalence with plain AdamW, and the
world_size=2 gloo plumbing test (BASELINE config #1) -- the reference has no tests
for its partitioning code at all."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zero_transformer_b200.models.gpt import GPTConfig, Transformer
from zero_transformer_b200.optim.schedule import WarmupCosine
from zero_transformer_b200.parallel.layout import FlatLayout
from zero_transformer_b200.parallel.zero1 import PlainAdamW, Zero1Engine

CFG = dict(embedding_dim=64, vocab_size=128, num_head=4, block_size=32, N=2, dropout=0.0)
LR = WarmupCosine(peak_value=1e-2, warmup_steps=2, decay_steps=50, end_value=1e-3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_layout_chunks_cover_everything():
    shapes = [("wte.weight", (128, 64)), ("blocks.0.ln1.weight", (64,)), ("blocks.0.attn.qkv.weight", (192, 64)),
              ("blocks.1.ln1.weight", (64,)), ("ln_f.weight", (70,))]
    L = FlatLayout(shapes, {}, world=4)
    # LayerNorm scales ride with the embed bucket; a block bucket holds the block's matrices
    assert [b.name for b in L.buckets] == ["embed", "block0"]
    assert [r.name for r in L.buckets[1].regions] == ["blocks.0.attn.qkv.weight"]
    assert L.total == L.shard_total * 4
    covered = torch.zeros(L.total)
    for r in L.regions:
        assert r.padded % (4 * L.align) == 0 and r.padded >= r.numel
        for k in range(4):
            s, e = r.chunk_range(k)
            covered[s:e] += 1
    assert torch.all(covered == 1)
    assert L.by_name["ln_f.weight"].decay is False and L.by_name["wte.weight"].decay is True
    flat = torch.arange(L.total, dtype=torch.float32)
    back = torch.zeros(L.total)
    for k in range(4):
        L.scatter_shard_to_flat(L.gather_shard_from_flat(flat, k), back, k)
    assert torch.equal(flat, back)


def _train(model, opt, batches, accum):
    losses = []
    for b in batches:
        opt.zero_grad()
        for mb in b.chunk(accum):
            loss = model.loss(mb, train=False)
            loss.backward()
            losses.append(loss.item())
        opt.step()
    return losses


@pytest.mark.parametrize("backend", ["nccl", "allreduce"])
def test_single_rank_engine_matches_plain_adamw(backend):
    torch.manual_seed(0)
    m1, m2 = Transformer(GPTConfig(**CFG)), Transformer(GPTConfig(**CFG))
    m2.load_state_dict(m1.state_dict())
    batches = [torch.randint(0, 128, (4, 32)) for _ in range(4)]
    plain = PlainAdamW(m1, lr_fn=LR, grad_accum_steps=2)
    eng = Zero1Engine(m2, lr_fn=LR, grad_accum_steps=2, backend=backend)
    l1 = _train(m1, plain, batches, 2)
    l2 = _train(m2, eng, batches, 2)
    assert l1 == pytest.approx(l2, rel=1e-5)
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6, msg=n)
    assert eng.step_count == 4 and eng.lr == pytest.approx(LR(4))


def _worker(rank, world, port, backend, q, clip_mode="elementwise", clip=1.0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = Transformer(GPTConfig(**CFG))
        eng = Zero1Engine(model, lr_fn=LR, grad_accum_steps=2, backend=backend, clip_mode=clip_mode, clip=clip)
        g = torch.Generator().manual_seed(1)
        batches = [torch.randint(0, 128, (4 * world, 32), generator=g) for _ in range(3)]
        for b in batches:
            mine = b.view(2, world, 2, 32)[:, rank].reshape(4, 32)  # rank's rows of each of the 2 micro-batches
            eng.zero_grad()
            for mb in mine.chunk(2):
                model.loss(mb, train=False).backward()
            eng.step()
        st = eng.consolidated_state(0)
        if rank == 0:
            torch.save({"params": {n: p.detach().clone() for n, p in model.named_parameters()}, "state": st}, q)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("backend,clip_mode,clip", [("nccl", "elementwise", 1.0), ("allreduce", "elementwise", 1.0),
                                                    ("nccl", "global_norm", 0.05)])
def test_world2_gloo_matches_single_process(backend, clip_mode, clip, tmp_path):
    """2 ranks x 2 micro-batches x 2 rows == 1 process x 2 micro-batches x 4 rows (both clipping modes; the
    global-norm threshold is small enough that the clip is active)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = str(tmp_path / "out.pt")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q, clip_mode, clip)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = torch.load(q, weights_only=False)
    params, state = res["params"], res["state"]

    torch.manual_seed(0)
    ref = Transformer(GPTConfig(**CFG))
    opt = PlainAdamW(ref, lr_fn=LR, grad_accum_steps=2, clip_mode=clip_mode, clip=clip)
    g = torch.Generator().manual_seed(1)
    batches = [torch.randint(0, 128, (4 * world, 32), generator=g) for _ in range(3)]
    for b in batches:
        opt.zero_grad()
        for mb in b.view(2, world * 2, 32):
            ref.loss(mb, train=False).backward()
        opt.step()
    for n, p in ref.named_parameters():
        torch.testing.assert_close(params[n], p.detach(), rtol=2e-5, atol=2e-6, msg=n)
        torch.testing.assert_close(state["params"][n], opt.master[n], rtol=2e-5, atol=2e-6, msg=n)
        torch.testing.assert_close(state["exp_avg"][n], opt.m[n], rtol=2e-4, atol=1e-7, msg=n)
    assert state["step"] == 3


def test_global_norm_clip_scales_the_whole_gradient():
    """clip_mode=global_norm: g <- g * min(1, clip / ||g||_2) over ALL parameters (non-default option)."""
    torch.manual_seed(0)
    m = Transformer(GPTConfig(**CFG))
    eng = Zero1Engine(m, lr_fn=LR, backend="nccl", clip_mode="global_norm", clip=1e-3)
    x = torch.randint(0, 128, (4, 32))
    eng.zero_grad()
    m.loss(x, train=False).backward()
    gnorm = torch.sqrt(sum(p.grad.float().pow(2).sum() for p in m.parameters()))
    eng.step()
    assert gnorm > 1e-3
    torch.testing.assert_close(eng.last_grad_norm.reshape(()), gnorm, rtol=1e-5, atol=0)
    # first Adam step: m = (1 - b1) * clipped g, so ||m|| = 0.1 * clip
    torch.testing.assert_close(eng.exp_avg.norm(), torch.tensor(0.1 * 1e-3), rtol=1e-3, atol=0)
    with pytest.raises(ValueError):
        Zero1Engine(Transformer(GPTConfig(**CFG)), lr_fn=LR, backend="nccl", clip_mode="nope")


def test_elementwise_clip_not_global_norm():
    """optax.clip(1.0) clamps each element (reference main_zero.py:161)."""
    from zero_transformer_b200.ops.reference import adamw_step_

    p, m, v = torch.zeros(3), torch.zeros(3), torch.zeros(3)
    g = torch.tensor([10.0, 0.5, -7.0])
    adamw_step_(p, g, m, v, lr=1.0, beta1=0.9, beta2=0.95, eps=1e-8, wd=0.0, step=1, clip=1.0)
    torch.testing.assert_close(m, torch.tensor([0.1, 0.05, -0.1]))


def test_partition_rules_drive_the_engine_layout():
    """The regex rule table (reference ``set_partitions_zero`` / ``create_opt_spec``) is what the engine builds its shard
    map from: buckets, decay flags and ownership all come out of ``parallel/partition.py``."""
    import pytest
    import torch

    from zero_transformer_b200.models.gpt import GPTConfig, Transformer
    from zero_transformer_b200.parallel import partition as P
    from zero_transformer_b200.parallel.zero1 import Zero1Engine

    m = Transformer(GPTConfig(embedding_dim=64, vocab_size=256, num_head=4, block_size=16, N=3))
    eng = Zero1Engine(m, lr_fn=lambda s: 1e-3, backend="nccl")
    ref = P.layout_for(m, 1)
    assert [(r.name, r.offset, r.decay, r.bucket) for r in eng.layout.regions] == [(r.name, r.offset, r.decay, r.bucket) for r in ref.regions]
    names = {b.name for b in eng.layout.buckets}
    assert names == {"embed", "block0", "block1", "block2"}
    rows = {r["name"]: r for r in P.describe(P.layout_for(m, 4))}
    assert rows["blocks.1.ln1.weight"]["bucket"] == "embed" and rows["blocks.1.ln1.weight"]["decay"] is False
    assert rows["blocks.2.mlp.fc_in.weight"]["bucket"] == "block2" and rows["blocks.2.mlp.fc_in.weight"]["decay"] is True
    assert rows["wte.weight"]["spec"] == ("dp", None) and rows["wte.weight"]["rows_per_rank"] == 64.0
    L4 = P.layout_for(m, 4)
    assert P.owner_of(L4, "wte.weight", 0) == 0 and P.owner_of(L4, "wte.weight", 256 * 64 - 1) == 3
    # custom rules: everything in one bucket, no decay on the embedding; an uncovered name is an error
    rules = (P.Rule(r"^wte", bucket="all", decay=False), P.Rule(r"^(blocks|ln_f)", bucket="all"))
    Lc = P.build_layout([(n, tuple(p.shape)) for n, p in m.named_parameters()], 2, rules=rules)
    assert len(Lc.buckets) == 1 and not next(r for r in Lc.regions if r.name == "wte.weight").decay
    with pytest.raises(KeyError):
        P.build_layout([("mystery.weight", (4, 4))], 2, rules=rules)
    assert P.main(["--model", "test", "--world", "2"]) == 0
