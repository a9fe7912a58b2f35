import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from zero_transformer_b200.ops import cuda as K
M, D = 16384, 2048
dy = (torch.randn(M, D, device="cuda") * 0.5).bfloat16(); w2 = (torch.randn(D, 4 * D, device="cuda") * 0.02).bfloat16(); z = torch.randn(M, 4 * D, device="cuda").bfloat16()
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
fl = 2.0 * M * D * 4 * D
g = t(lambda: K._C.gemm(dy, w2, epilogue=K.EPI_GELU_BWD, aux_in=z, b_mn=True, two_cta=True)); p = t(lambda: K._C.gemm(dy, w2, b_mn=True, two_cta=True))
out = K._C.gemm(dy[:2048], w2, epilogue=K.EPI_GELU_BWD, aux_in=z[:2048], b_mn=True, two_cta=True).float()
zf = z[:2048].float().requires_grad_(True); torch.nn.functional.gelu(zf, approximate="tanh").backward(dy[:2048].float() @ w2.float())
err = (out - zf.grad).abs().max().item(); ref = zf.grad.abs().max().item()
print(json.dumps({"dgrad_gelu_bwd_tflops": round(fl / g / 1e9, 1), "dgrad_plain_tflops": round(fl / p / 1e9, 1), "ratio": round(p / g, 3), "max_abs_err": err, "ref_max": ref}))
