"""Reduce-scatter / all-gather bus-bandwidth sweep and fused-path roofline fractions (BASELINE config #5).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 bench/comm_bench.py

Measures, on the GPT-2 1.3B parameter layout (2.62 GB of bf16 gradients / parameters):
  * NCCL reduce_scatter_tensor / all_gather_into_tensor on the same flat buffers (baseline),
  * our P2P kernels: standalone reduce-scatter, all-gather pull,
  * the fused reduce-scatter + AdamW kernel (per-step total over all buckets),
  * the all-gather-fused QKV GEMM vs. the plain QKV GEMM (one block bucket).
All times are CUDA-event times, max over ranks.  "busbw" follows the NCCL convention
bytes * (n-1)/n / time.  Roofline for a fused kernel = max(HBM time, NVLink time, math time) with the
measured denominators (HBM copy bandwidth from MEASURED_PEAKS.json, 770 GB/s per direction NVLink,
cuBLAS bf16 peak).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from zero_transformer_b200.models.gpt import GPTConfig
from zero_transformer_b200.ops import cuda as K
from zero_transformer_b200.parallel import dist as pdist
from zero_transformer_b200.parallel.fused import FLAG_SHARD_READY, FusedComm, pack_segments
from zero_transformer_b200.parallel.layout import FlatLayout
from zero_transformer_b200.utils.metrics import measured_peaks

NVLINK_GBS = 770.0  # measured peer-copy bandwidth per direction per GPU (profiling guide)


def timed(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    return pdist.max_over_ranks(t, "cuda")


def main():
    env = pdist.setup_dp()
    n, r, dev = env.world, env.rank, env.device
    pk = measured_peaks()
    cfg = GPTConfig(embedding_dim=2048, vocab_size=50304, num_head=16, block_size=1024, N=24)
    D, V = cfg.embedding_dim, cfg.vocab_size
    shapes = [("wte.weight", (V, D))]
    for i in range(cfg.N):
        shapes += [(f"blocks.{i}.ln1.weight", (D,)), (f"blocks.{i}.attn.qkv.weight", (3 * D, D)), (f"blocks.{i}.attn.out.weight", (D, D)),
                   (f"blocks.{i}.ln2.weight", (D,)), (f"blocks.{i}.mlp.fc_in.weight", (4 * D, D)), (f"blocks.{i}.mlp.fc_out.weight", (D, 4 * D))]
    shapes.append(("ln_f.weight", (D,)))
    L = FlatLayout(shapes, {}, n)
    fc = FusedComm(L, r, n, None, dev, torch.bfloat16, torch.bfloat16)
    C = fc._C
    fc.flat_grad.normal_()
    fc.flat_param.normal_()
    total_bytes = L.total * 2
    out = {"world": n, "params": L.total, "bytes_bf16": total_bytes, "nvls": bool(fc.nvls)}
    wire = total_bytes * (n - 1) / n  # bytes each rank must receive

    master = torch.randn(L.shard_total, device=dev)
    m, v = torch.zeros_like(master), torch.zeros_like(master)
    kw = dict(master=master, m=m, v=v, lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, wd=0.1, step=1, clip=1.0, grad_scale=1.0 / n)

    def fused_rs_adam():
        fc.begin_step()
        for b in range(fc.nb):
            fc.bucket_update(b, **kw)
        torch.cuda.current_stream().wait_stream(fc.stream)

    t = timed(fused_rs_adam)
    hbm = (2 * L.shard_total * n + 12 * L.shard_total + 14 * L.shard_total) / (pk["hbm_gbs"] * 1e9) * 1e3  # local part of the reads + state
    nv = wire / (NVLINK_GBS * 1e9) * 1e3
    if fc.nvls:   # switch-side reduction: the shard comes in once; the push sends the shard out once, (n-1)/n of the model comes in
        nv = max(2 * L.shard_total, wire) / (NVLINK_GBS * 1e9) * 1e3
    out["rs_adam_fused"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1), "roofline_ms": round(max(hbm, nv), 3),
                            "roofline_frac": round(max(hbm, nv) / t, 3), "bound": "nvlink" if nv > hbm else "hbm"}

    if n > 1:
        def ag_pull():
            fc.begin_step()   # (_pushed reset: measures the P2P pull even when the NVLS push is available)
            with torch.cuda.stream(fc.stream):
                for b in range(fc.nb):
                    fc._signal(FLAG_SHARD_READY, b, fc.epoch)
            fc._pending = [True] * fc.nb
            fc.prefetch_all()
            fc.wait_all()

        t = timed(ag_pull)
        out["allgather_pull"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1), "roofline_ms": round(nv, 3), "roofline_frac": round(nv / t, 3)}

        # standalone P2P reduce-scatter (whole model in one launch)
        segs, tiles = pack_segments([(rg.offset + r * rg.chunk, rg.shard_offset, rg.chunk, 0) for rg in L.regions], dev, 2048)
        rs_out = torch.empty(L.shard_total, dtype=torch.bfloat16, device=dev)

        def rs_p2p():
            C.reduce_scatter(rs_out, fc.grad_ptrs, segs, len(L.regions), tiles, 0, 0)

        dist.barrier()
        t = timed(rs_p2p)
        out["reduce_scatter_p2p"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1), "roofline_frac": round(nv / t, 3)}

        # NCCL on equally sized contiguous buffers
        g_in = torch.randn(L.total, device=dev).bfloat16()
        g_out = torch.empty(L.total // n, dtype=torch.bfloat16, device=dev)
        t = timed(lambda: dist.reduce_scatter_tensor(g_out, g_in))
        out["nccl_reduce_scatter"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1)}
        t = timed(lambda: dist.all_gather_into_tensor(g_in, g_out))
        out["nccl_all_gather"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1)}
        t = timed(lambda: dist.all_reduce(g_in))
        out["nccl_all_reduce"] = {"ms": round(t, 3), "busbw_gbs": round(2 * wire / t / 1e6, 1)}

        # all-gather fused into the QKV GEMM of one block vs. plain GEMM (+ what a separate gather costs)
        Mtok = 16384
        x = torch.randn(Mtok, D, device=dev).bfloat16()
        b = next(bk.index for bk in L.buckets if bk.name == "block0")
        name = "blocks.0.attn.qkv.weight"
        w = L.view(fc.flat_param, name)
        bucket_bytes = sum(rg.padded for rg in L.buckets[b].regions) * 2
        bwire = bucket_bytes * (n - 1) / n

        def ag_gemm():
            fc.begin_step()
            with torch.cuda.stream(fc.stream):
                fc._signal(FLAG_SHARD_READY, b, fc.epoch)
            torch.cuda.current_stream().wait_stream(fc.stream)
            fc.gemm_ag(b, name, x, w)

        t_ag = timed(ag_gemm)
        t_plain = timed(lambda: K._C.gemm(x, w, two_cta=False))
        t_plain2 = timed(lambda: K._C.gemm(x, w, two_cta=True))
        segs_b, nb_, tiles_b = fc.pull[b]
        t_pull = timed(lambda: C.allgather_pull(fc.flat_param, fc._param.ptrs, segs_b, nb_, tiles_b, 0, 0))
        fl = 2.0 * Mtok * 3 * D * D
        math_ms = fl / (pk["bf16_tflops"] * 1e12) * 1e3
        nv_ms = bwire / (NVLINK_GBS * 1e9) * 1e3
        out["ag_gemm_fused"] = {"ms": round(t_ag, 3), "plain_gemm_1cta_ms": round(t_plain, 3), "plain_gemm_2cta_ms": round(t_plain2, 3),
                                "separate_pull_ms": round(t_pull, 3), "bucket_mb": round(bucket_bytes / 1e6, 1),
                                "roofline_ms": round(max(math_ms, nv_ms), 3), "roofline_frac": round(max(math_ms, nv_ms) / t_ag, 3),
                                "bound": "nvlink" if nv_ms > math_ms else "math"}
    if env.is_main:
        print(json.dumps(out))
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/comm_bench_n{n}.json", "w"), indent=1)
    pdist.shutdown()


if __name__ == "__main__":
    main()
