"""Per-phase cycle breakdown of one softmax thread in the attention backward kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zero_transformer_b200.ops import cuda as K
from zero_transformer_b200.ops.reference import get_slopes
B, T, H, hd = 16, 1024, 16, 128
qkv = torch.randn(B, T, 3 * H * hd, device="cuda").bfloat16()
slopes = torch.tensor(get_slopes(H), device="cuda")
rng = K.rng_state("cuda")
out, lse = K._C.attention_fwd(qkv, slopes, H, hd ** -0.5, 0.1, rng, 1)
dout = torch.randn_like(out)
for _ in range(2):
    K._C.attention_bwd(qkv, out, dout, lse, slopes, H, hd ** -0.5, 0.1, rng, 1)
dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
K._C.attention_bwd(qkv, out, dout, lse, slopes, H, hd ** -0.5, 0.1, rng, 1, dbg=dbg)
torch.cuda.synchronize()
d = dbg.tolist()
n = max(d[6], 1)
names = ["wait S", "P phase", "wait dP", "dS phase", "wait dV+dQ", "dQ drain"]
tot = sum(d[:6])
print(f"iterations sampled: {n}; mean cycles per iteration: {tot / n:.0f}")
for nm, v in zip(names, d[:6]):
    print(f"  {nm:12s} {v / n:9.0f} cycles  {100 * v / tot:5.1f} %")
