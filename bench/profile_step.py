"""Per-kernel time breakdown of one training step (torch.profiler / CUPTI), written to
gpurun_out/step_profile.json + a text table.  Not a timing source for bench numbers."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# per-kernel durations: with programmatic dependent launch a kernel's CUPTI interval starts when it is scheduled, i.e. it
# includes the time it spends in griddepcontrol.wait behind its predecessor -- profile with plain stream order
os.environ.setdefault("ZTB_PDL", "0")
import torch
from torch.profiler import ProfilerActivity, profile

from zero_transformer_b200.config import DEFAULT_MODEL_CFG, load_config
from zero_transformer_b200.parallel import dist as pdist
from zero_transformer_b200.train import Trainer

model = os.environ.get("ZTB_BENCH_MODEL", "1_3b")
rows, accum, seq = int(os.environ.get("ZTB_BENCH_ROWS", "32")), int(os.environ.get("ZTB_BENCH_ACCUM", "2")), 1024
env = pdist.setup_dp()
cfg = load_config(None, [f"model.size={model}", f"training.batch_size={rows}", f"training.train_context={seq}", f"data.max_context={seq}",
                         f"training.gradient_accumulation_steps={accum}", "training.evaluation_frequency=0",
                         "data.checkpoint_directory=gpurun_out/prof_ckpt"])
tr = Trainer(cfg, DEFAULT_MODEL_CFG, env, log_metrics=False)
x = torch.randint(0, 50304, (rows, seq), dtype=torch.int32, device=env.device)
for _ in range(3):
    tr.train_step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.train_step(x)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_step(x)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in ev) / 1e3
out = {"model": model, "step_ms_unprofiled": step_ms, "sum_kernel_ms": total,
       "kernels": [{"name": e.key[:110], "calls": e.count, "total_ms": round(e.device_time_total / 1e3, 3),
                    "avg_us": round(e.device_time_total / max(e.count, 1), 1), "pct": round(100 * e.device_time_total / 1e3 / total, 2)} for e in ev[:40]]}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/step_profile.json", "w"), indent=1)
print(f"step {step_ms:.1f} ms unprofiled; sum of kernel time {total:.1f} ms")
for k in out["kernels"][:28]:
    print(f"{k['pct']:6.2f}% {k['total_ms']:9.3f} ms {k['calls']:5d}x {k['avg_us']:9.1f} us  {k['name']}")
