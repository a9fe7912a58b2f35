"""Kernel table of ONE training step on every rank (torch.profiler / CUPTI) -- evidence that the ZeRO-1 data plane of the
fused backend launches no NCCL kernel between step boundaries.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 bench/kernel_table.py

Rank 0 writes gpurun_out/kernel_table_n<N>.json / .md: per kernel name the launch count and total device time, grouped
into compute / zero1 (adamw_shard, signal, wait_flags, allgather_pull, gemm with all-gather) / nccl / other, plus the
NCCL kernel count seen on ANY rank (must be 0).  Diagnostic only; not a timing source."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# per-kernel durations: with programmatic dependent launch a kernel's CUPTI interval starts when it is scheduled, i.e. it
# includes the time it spends in griddepcontrol.wait behind its predecessor -- profile with plain stream order
os.environ.setdefault("ZTB_PDL", "0")
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

from zero_transformer_b200.config import DEFAULT_MODEL_CFG, load_config
from zero_transformer_b200.parallel import dist as pdist
from zero_transformer_b200.train import Trainer

model = os.environ.get("ZTB_BENCH_MODEL", "1_3b")
rows, accum, seq = int(os.environ.get("ZTB_BENCH_ROWS", "32")), int(os.environ.get("ZTB_BENCH_ACCUM", "2")), int(os.environ.get("ZTB_BENCH_SEQ", "1024"))
env = pdist.setup_dp()
cfg = load_config(None, [f"model.size={model}", f"training.batch_size={rows * env.world}", f"training.train_context={seq}", f"data.max_context={seq}",
                         f"training.gradient_accumulation_steps={accum}", "training.evaluation_frequency=0", "training.grad_dtype=bfloat16",
                         "data.checkpoint_directory=gpurun_out/prof_ckpt"])
tr = Trainer(cfg, DEFAULT_MODEL_CFG, env, log_metrics=False)
x = torch.randint(0, 50304, (rows, seq), dtype=torch.int32, device=env.device)
for _ in range(3):
    tr.train_step(x)
torch.cuda.synchronize()
if env.world > 1:
    dist.barrier()          # the only NCCL kernel of this script: outside the profiled step
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(x)
    tr.train_step(x)        # second step: the forward that consumes the previous step's all-gather
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
rowsout = []
for e in ev:
    n = e.key
    low = n.lower()
    grp = ("nccl" if "nccl" in low else
           "zero1" if any(k in n for k in ("adamw_shard", "signal_kernel", "wait_flags", "allgather_pull", "reduce_scatter_kernel",
                                           "gemm2_tcgen05_kernel<false, false, 0, true>", "gemm_tcgen05_kernel<false, false, true>")) else
           "compute" if "ztb::" in n else "other")
    rowsout.append({"name": n[:120], "group": grp, "calls": e.count, "total_ms": round(e.device_time_total / 1e3, 3)})
nccl_here = sum(r["calls"] for r in rowsout if r["group"] == "nccl")
t = torch.tensor([nccl_here], device=env.device, dtype=torch.int64)
if env.world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if env.is_main:
    rowsout.sort(key=lambda r: -r["total_ms"])
    out = {"world": env.world, "model": model, "steps_profiled": 2, "nvls": bool(getattr(tr.engine._fused, "nvls", False)),
           "nccl_kernels_between_step_boundaries_max_over_ranks": int(t.item()), "kernels": rowsout}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/kernel_table_n{env.world}.json", "w"), indent=1)
    with open(f"gpurun_out/kernel_table_n{env.world}.md", "w") as f:
        f.write(f"# Kernel table, 2 profiled steps, world {env.world}, data plane {'NVLS' if out['nvls'] else 'P2P'}\n\n"
                f"NCCL kernels between step boundaries (max over ranks): **{out['nccl_kernels_between_step_boundaries_max_over_ranks']}**\n\n"
                "| group | calls | total ms | kernel |\n|---|---|---|---|\n")
        for r in rowsout[:40]:
            f.write(f"| {r['group']} | {r['calls']} | {r['total_ms']} | `{r['name'][:90]}` |\n")
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
    for r in rowsout[:12]:
        print(r)
tr.close()
pdist.shutdown()
