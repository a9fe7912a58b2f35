"""Reduce-scatter / all-gather bus-bandwidth sweep and fused-path roofline fractions (BASELINE config #5).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 bench/comm_bench.py

Measures, on the GPT-2 1.3B parameter layout (2.62 GB of bf16 gradients / parameters):
  * NCCL reduce_scatter_tensor / all_gather_into_tensor on the same flat buffers (baseline),
  * our P2P kernels: standalone reduce-scatter, all-gather pull,
  * the fused reduce-scatter + AdamW kernel (per-step total over all buckets),
  * the all-gather-fused QKV GEMM vs. the plain QKV GEMM (one block bucket).
All times are CUDA-event times, max over ranks.  "busbw" follows the NCCL convention
bytes * (n-1)/n / time.

ONE roofline per fused kernel (round-1 reported two inconsistent ones): the slower of
  link   bytes that must cross this GPU's NVLink port in the busier direction / link bandwidth
  hbm    bytes this GPU's HBM must move / measured copy bandwidth (MEASURED_PEAKS.json)
  math   FLOPs / measured cuBLAS bf16 burst rate (GEMM-fused kernels only)
with P = parameter count, n = ranks, bf16 gradients and parameters, fp32 master + 2 moments:
  fused RS+AdamW+AG, NVLS plane   link: 2P out (the switch pulls every rank's copy of every shard) + 2P/n out (push of
                                  the new shard); 2P/n + 2P in  ->  2P (1 + 1/n) each way
                                  hbm : 2P (gradient reads) + 24 P/n (master, m, v read + write) + 2P (gathered params written)
  fused RS+AdamW, P2P plane       link: 2P (n-1)/n in (peer gradient chunks); hbm: 2P + 24 P/n + 2P/n
  all-gather pull / AG in GEMM    link: 2 B (n-1)/n in for a bucket of B parameters
The fraction is reported against the MEASURED link bandwidth (770 GB/s per direction, profiling guide) and, for
context, against the nominal 900 GB/s.
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
    only_ag = os.environ.get("COMM_ONLY_AG") == "1"   # quick mode: only the all-gather GEMM section
    T = (lambda fn, **kw: float("nan")) if only_ag else timed
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

    P = L.total
    have_mc = bool(fc.nvls)

    def rs_adam_record(nvls_plane):
        fc.nvls = nvls_plane
        t = T(fused_rs_adam)
        if nvls_plane:
            link_bytes = 2 * P * (1 + 1 / n)
            hbm_bytes = 2 * P + 24 * P / n + 2 * P
        else:
            link_bytes = 2 * P * (n - 1) / n
            hbm_bytes = 2 * P + 24 * P / n + 2 * P / n
        hbm = hbm_bytes / (pk["hbm_gbs"] * 1e9) * 1e3
        nv = link_bytes / (NVLINK_GBS * 1e9) * 1e3
        nv900 = link_bytes / 900e9 * 1e3
        return {"ms": round(t, 3), "plane": "nvls" if nvls_plane else "p2p", "includes_allgather": bool(nvls_plane),
                "link_bytes_per_rank_per_direction": int(link_bytes), "hbm_bytes_per_rank": int(hbm_bytes),
                "link_ms_at_770": round(nv, 3), "link_ms_at_900": round(nv900, 3), "hbm_ms": round(hbm, 3),
                "roofline_ms": round(max(hbm, nv), 3), "roofline_frac": round(max(hbm, nv) / t, 3),
                "roofline_frac_vs_900": round(max(hbm, nv900) / t, 3), "bound": "nvlink" if nv > hbm else "hbm",
                "busbw_gbs": round(wire / t / 1e6, 1)}

    out["rs_adam_fused"] = rs_adam_record(have_mc)            # the default data plane
    if have_mc:
        out["rs_adam_fused_p2p"] = rs_adam_record(False)       # same kernel, peer loads instead of the switch reduction
        fc.nvls = True
    nv = wire / (NVLINK_GBS * 1e9) * 1e3   # plain reduce-scatter / all-gather rows below: 2P (n-1)/n over the link

    if n > 1:
        def ag_pull():
            fc.begin_step()   # (_pushed reset: measures the P2P pull even when the NVLS push is available)
            with torch.cuda.stream(fc.stream):
                for b in range(fc.nb):
                    fc._signal(FLAG_SHARD_READY, b, fc.epoch)
            fc._pending = [True] * fc.nb
            fc.prefetch_all()
            fc.wait_all()

        t = T(ag_pull)
        out["allgather_pull"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1), "roofline_ms": round(nv, 3), "roofline_frac": round(nv / t, 3)}

        # standalone P2P reduce-scatter (whole model in one launch)
        segs, tiles = pack_segments([(rg.offset + r * rg.chunk, rg.shard_offset, rg.chunk, 0) for rg in L.regions], dev, 2048)
        rs_out = torch.empty(L.shard_total, dtype=torch.bfloat16, device=dev)

        def rs_p2p():
            C.reduce_scatter(rs_out, fc.grad_ptrs, segs, len(L.regions), tiles, 0, 0)

        dist.barrier()
        t = T(rs_p2p)
        out["reduce_scatter_p2p"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1), "roofline_frac": round(nv / t, 3)}

        # NCCL on equally sized contiguous buffers
        g_in = torch.randn(L.total, device=dev).bfloat16()
        g_out = torch.empty(L.total // n, dtype=torch.bfloat16, device=dev)
        t = T(lambda: dist.reduce_scatter_tensor(g_out, g_in))
        out["nccl_reduce_scatter"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1)}
        t = T(lambda: dist.all_gather_into_tensor(g_in, g_out))
        out["nccl_all_gather"] = {"ms": round(t, 3), "busbw_gbs": round(wire / t / 1e6, 1)}
        t = T(lambda: dist.all_reduce(g_in))
        out["nccl_all_reduce"] = {"ms": round(t, 3), "busbw_gbs": round(2 * wire / t / 1e6, 1)}

        # all-gather fused into the QKV GEMM of one block vs. plain GEMM (+ what a separate gather costs)
        Mtok = 16384
        x = torch.randn(Mtok, D, device=dev).bfloat16()
        b = next(bk.index for bk in L.buckets if bk.name == "block0")
        name = "blocks.0.attn.qkv.weight"
        w = L.view(fc.flat_param, name)
        bucket_bytes = sum(rg.padded for rg in L.buckets[b].regions) * 2
        bwire = bucket_bytes * (n - 1) / n

        # publish the bucket once; the timed region is the GEMM launch alone, like the plain GEMM it is compared with (in the
        # training step the SHARD_READY signal belongs to the update kernel's stream, not to the GEMM)
        fc.begin_step()
        with torch.cuda.stream(fc.stream):
            fc._signal(FLAG_SHARD_READY, b, fc.epoch)
        torch.cuda.current_stream().wait_stream(fc.stream)
        torch.cuda.synchronize()
        dist.barrier()

        def ag_gemm():
            fc.gemm_ag(b, name, x, w)

        fc.ag_two_cta = False
        t_ag1 = timed(ag_gemm)
        fc.ag_two_cta = True
        t_ag = timed(ag_gemm)
        t_plain = timed(lambda: K._C.gemm(x, w, two_cta=False))
        t_plain2 = timed(lambda: K._C.gemm(x, w, two_cta=True))
        segs_b, nb_, tiles_b = fc.pull[b]
        t_pull = timed(lambda: C.allgather_pull(fc.flat_param, fc._param.ptrs, segs_b, nb_, tiles_b, 0, 0))
        fl = 2.0 * Mtok * 3 * D * D
        math_ms = fl / (pk["bf16_tflops"] * 1e12) * 1e3
        nv_ms = bwire / (NVLINK_GBS * 1e9) * 1e3
        out["ag_gemm_fused"] = {"ms": round(t_ag, 3), "kernel": "cta-pair + bulk-copy gather ring", "ms_1cta_register_gather": round(t_ag1, 3),
                                "vs_plain_2cta": round(t_ag / t_plain2, 3), "plain_gemm_1cta_ms": round(t_plain, 3), "plain_gemm_2cta_ms": round(t_plain2, 3),
                                "separate_pull_ms": round(t_pull, 3), "bucket_mb": round(bucket_bytes / 1e6, 1),
                                "roofline_ms": round(max(math_ms, nv_ms), 3), "roofline_frac": round(max(math_ms, nv_ms) / t_ag, 3),
                                "bound": "nvlink" if nv_ms > math_ms else "math", "variant": os.environ.get("ZTB_AG_VARIANT", "1")}
    if env.is_main:
        print(json.dumps(out))
        os.makedirs("gpurun_out", exist_ok=True)
        if not only_ag:
            json.dump(out, open(f"gpurun_out/comm_bench_n{n}.json", "w"), indent=1)
    pdist.shutdown()


if __name__ == "__main__":
    main()
