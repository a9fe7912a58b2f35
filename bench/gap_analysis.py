"""Where does the main stream idle?  Profiles one training step (CUPTI via torch.profiler), then lists the
idle gaps between consecutive kernels of the busiest stream grouped by (previous kernel -> next kernel).
Diagnostic only; not a timing source."""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ZTB_PDL", "0")   # with PDL a kernel's interval starts before its predecessor ends: gaps would be negative
import torch
from torch.profiler import ProfilerActivity, profile

from zero_transformer_b200.config import DEFAULT_MODEL_CFG, load_config
from zero_transformer_b200.parallel import dist as pdist
from zero_transformer_b200.train import Trainer

model = os.environ.get("ZTB_BENCH_MODEL", "1_3b")
rows, accum, seq = int(os.environ.get("ZTB_BENCH_ROWS", "32")), int(os.environ.get("ZTB_BENCH_ACCUM", "2")), 1024
env = pdist.setup_dp()
cfg = load_config(None, [f"model.size={model}", f"training.batch_size={rows}", f"training.train_context={seq}", f"data.max_context={seq}",
                         f"training.gradient_accumulation_steps={accum}", "training.evaluation_frequency=0", "training.grad_dtype=bfloat16",
                         "data.checkpoint_directory=gpurun_out/prof_ckpt"])
tr = Trainer(cfg, DEFAULT_MODEL_CFG, env, log_metrics=False)
x = torch.randint(0, 50304, (rows, seq), dtype=torch.int32, device=env.device)
for _ in range(3):
    tr.train_step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_step(x)
    tr.train_step(x)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/step_trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
by_stream = collections.defaultdict(list)
for e in ks:
    by_stream[e["args"].get("stream")].append(e)
main = max(by_stream.values(), key=lambda v: sum(e["dur"] for e in v))
main.sort(key=lambda e: e["ts"])
span = main[-1]["ts"] + main[-1]["dur"] - main[0]["ts"]
busy = sum(e["dur"] for e in main)
print(f"streams: { {s: (len(v), round(sum(e['dur'] for e in v) / 1e3, 1)) for s, v in by_stream.items()} }")
print(f"main stream: {len(main)} kernels over 2 steps, span {span / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms, idle {(span - busy) / 1e3:.1f} ms")
gaps = collections.defaultdict(lambda: [0, 0.0])
hist = collections.Counter()
short = lambda n: n.replace("void ", "").replace("ztb::", "").split("(")[0][:60]
for a, b in zip(main, main[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    k = (short(a["name"]), short(b["name"]))
    gaps[k][0] += 1
    gaps[k][1] += g
    hist[min(int(g // 2) * 2, 40)] += 1
print("gap histogram (us bucket: count):", sorted(hist.items()))
for k, (n, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t / 1e3:8.2f} ms  {n:5d}x  avg {t / n:7.1f} us   {k[0]}  ->  {k[1]}")
