"""tcgen05 attention vs flash_attn library (fwd), CUDA-event timed."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from zero_transformer_b200.ops import cuda as K
from zero_transformer_b200.ops.reference import get_slopes


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


B, T, H, hd = 16, 1024, 16, 128
qkv = torch.randn(B, T, 3 * H * hd, device="cuda").bfloat16()
slopes = torch.tensor(get_slopes(H), device="cuda")
rng = K.rng_state("cuda")
fl = 4 * B * H * T * T * hd / 2
for p in (0.0, 0.1):
    t = timeit(lambda: K._C.attention_fwd(qkv, slopes, H, hd ** -0.5, p, rng, 1))
    print(json.dumps({"kernel": "ztb_attn_fwd", "dropout": p, "ms": round(t, 4), "tflops_causal": round(fl / t / 1e9, 1)}))
from flash_attn import flash_attn_qkvpacked_func

q5 = qkv.view(B, T, 3, H, hd)
t = timeit(lambda: flash_attn_qkvpacked_func(q5, 0.0, hd ** -0.5, True, alibi_slopes=slopes))
print(json.dumps({"kernel": "flash_attn2_fwd", "ms": round(t, 4), "tflops_causal": round(fl / t / 1e9, 1)}))

from zero_transformer_b200.ops.attention_cuda import alibi_attention

qg = qkv.clone().requires_grad_(True)
dout = torch.randn(B, T, H * hd, device="cuda").bfloat16()
for dq16 in (False, True):
    K.CONFIG.attn_dq_bf16 = dq16
    for p in (0.0, 0.1):
        K.begin_forward("cuda")
        o = alibi_attention(qg, slopes, H, hd ** -0.5, p)
        t = timeit(lambda: o.backward(dout, retain_graph=True))
        print(json.dumps({"kernel": "ztb_attn_bwd(+delta" + ("+bf16 dQ reduce" if dq16 else "+fp32 dQ scratch+convert") + ")", "dropout": p,
                          "ms": round(t, 4), "tflops_causal": round(2.5 * fl / t / 1e9, 1)}))
K.CONFIG.attn_dq_bf16 = False
of = flash_attn_qkvpacked_func(qg.view(B, T, 3, H, hd), 0.0, hd ** -0.5, True, alibi_slopes=slopes)
t = timeit(lambda: of.backward(dout.view(B, T, H, hd), retain_graph=True))
print(json.dumps({"kernel": "flash_attn2_bwd", "ms": round(t, 4), "tflops_causal": round(2.5 * fl / t / 1e9, 1)}))
