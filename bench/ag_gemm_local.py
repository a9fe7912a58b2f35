"""All-gather-fused QKV GEMM on ONE GPU with a fake world: the "peers" are other buffers on the same device.

Exercises exactly the kernel the P2P data plane runs (pull list, gather ring, per-owner done counters, owned-column-block
first raster) without needing a multi-GPU box, so its overhead over the plain CTA-pair GEMM can be profiled with ncu and
iterated on cheaply.  The NVLink part (latency, 770 GB/s) is of course absent: bench/comm_bench.py measures that at n = 2/4/8.

    python bench/ag_gemm_local.py            # FAKE_WORLD=2|4|8
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from zero_transformer_b200.ops import cuda as K
from zero_transformer_b200.parallel.fused import AG2_TILE, COPY_TILE, MAX_RANKS, pack_segments
from zero_transformer_b200.parallel.layout import FlatLayout


def timed(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    D, M = int(os.environ.get("D", 2048)), int(os.environ.get("M", 16384))
    for W in [int(w) for w in os.environ.get("FAKE_WORLD", "2,8").split(",")]:
        shapes = [("blocks.0.attn.qkv.weight", (3 * D, D)), ("blocks.0.attn.out.weight", (D, D)),
                  ("blocks.0.mlp.fc_in.weight", (4 * D, D)), ("blocks.0.mlp.fc_out.weight", (D, 4 * D))]
        L = FlatLayout(shapes, {}, W)
        full = (torch.randn(L.total, device=dev) * 0.05).bfloat16()
        local = full.clone()
        peers = [local] + [full.clone() for _ in range(W - 1)]
        r0 = L.regions[0]
        # rank 0 owns chunk 0 of every region; wipe the rest of its copy
        for rg in L.regions:
            local[rg.offset + rg.chunk: rg.offset + rg.padded].zero_()
        x = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
        w = local[r0.offset: r0.offset + r0.numel].view(3 * D, D)
        wfull = full[r0.offset: r0.offset + r0.numel].view(3 * D, D)
        flags = torch.ones(MAX_RANKS, dtype=torch.int32, device=dev)
        done = torch.zeros(8, dtype=torch.int32, device=dev)
        order = list(range(1, W))
        rec = {"fake_world": W, "M": M, "D": D, "bucket_mb": round(L.total * 2 / 1e6, 1)}
        for two_cta, tile in ((True, AG2_TILE), (False, COPY_TILE)):
            pl = [(r0.offset + p * r0.chunk, 0, r0.chunk, p) for p in order]
            n_b = len(pl)
            pl += [(rg.offset + p * rg.chunk, 0, rg.chunk, p) for p in order for rg in L.regions[1:]]
            segs, tiles = pack_segments(pl, dev, tile)
            args = (x, w, [t.data_ptr() for t in peers], local, segs, len(pl), n_b, tiles, flags.data_ptr(), 1, done,
                    (r0.chunk + tile - 1) // tile, r0.chunk // D, W, 0, two_cta, 0)
            out = K._C.gemm_ag(*args)
            torch.cuda.synchronize()
            ref = x.float() @ wfull.float().t()
            err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
            assert torch.equal(local, full), "gather incomplete"
            assert err < 2e-2, err
            done.zero_()
            state = {"base": 0}
            b_tiles = args[11]

            def run_ag(a=args):   # cumulative landing counters: no memset launch in front of the GEMM (what FusedComm.gemm_ag does)
                out = K._C.gemm_ag(*a, done_base=state["base"])
                state["base"] = (state["base"] + b_tiles) & 0xFFFFFFFF
                return out

            rec["ag_2cta_ms" if two_cta else "ag_1cta_ms"] = round(timed(run_ag), 4)
            assert torch.equal(run_ag(), out), "cumulative-counter launch differs"
            if two_cta:   # a 256-row GEMM: the kernel time is the gather's own duration (driver-thread latency chain x tiles per CTA)
                xs = x[:256].contiguous()
                args_s = (xs,) + args[1:]
                rec["gather_only_ms"] = round(timed(lambda: run_ag(args_s)), 4)
                rec["gather_tiles_per_cta"] = round(tiles / 148, 1)
        rec["plain_2cta_ms"] = round(timed(lambda: K._C.gemm(x, wfull, two_cta=True)), 4)
        rec["plain_2cta_raster_n_ms"] = round(timed(lambda: K._C.gemm(x, wfull, two_cta=True, raster_n=1)), 4)
        rec["plain_1cta_ms"] = round(timed(lambda: K._C.gemm(x, wfull, two_cta=False)), 4)
        rec["vs_plain_2cta"] = round(rec["ag_2cta_ms"] / rec["plain_2cta_ms"], 3)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
