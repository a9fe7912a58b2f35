"""Weight-gradient GEMM (K = tokens): whole-tile waves vs. K-split last wave (gemm2_tcgen05.cu), plus a
numerics check of the split path against an fp32 reference."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench.gemm_bench import timeit
from zero_transformer_b200.ops import cuda as K

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = 2048
for name, n_out, n_in in [("qkv", 3 * D, D), ("out", D, D), ("fc_in", 4 * D, D), ("fc_out", D, 4 * D)]:
    dy = torch.randn(M, n_out, device="cuda").bfloat16()
    x = torch.randn(M, n_in, device="cuda").bfloat16()
    fl = 2.0 * M * n_out * n_in
    ref = dy.float().t() @ x.float()
    row = {"name": name, "tiles": (n_out // 256) * (n_in // 256)}
    for acc in (False, True):
        for ns in (True, False):
            c = torch.zeros(n_out, n_in, device="cuda", dtype=torch.bfloat16)
            K._C.gemm(dy, x, C=c, a_mn=True, b_mn=True, two_cta=True, accumulate=acc, no_split=ns)
            err = ((c.float() - ref).norm() / ref.norm()).item()
            t = timeit(lambda: K._C.gemm(dy, x, C=c, a_mn=True, b_mn=True, two_cta=True, accumulate=acc, no_split=ns))
            row[f"{'acc' if acc else 'store'}_{'whole' if ns else 'split'}"] = {"ms": round(t, 4), "tflops": round(fl / t / 1e9, 1), "rel_err": round(err, 5)}
    t = timeit(lambda: dy.t() @ x)
    row["cublas_tflops"] = round(fl / t / 1e9, 1)
    print(json.dumps(row), flush=True)
