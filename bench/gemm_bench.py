"""tcgen05 GEMM vs cuBLAS on the model's shapes (CUDA-event timed, L2 flushed between iterations)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from zero_transformer_b200.ops import cuda as K


def timeit(fn, iters=10, warmup=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    D = 2048
    shapes = [("qkv", M, 3 * D, D), ("out", M, D, D), ("fc_in", M, 4 * D, D), ("fc_out", M, D, 4 * D), ("lm_head", M, 50304, D)]
    rows = []
    for name, m, n, k in shapes:
        a = torch.randn(m, k, device="cuda").bfloat16()
        b = torch.randn(n, k, device="cuda").bfloat16()
        fl = 2.0 * m * n * k
        t_tc = timeit(lambda: K._C.gemm(a, b, two_cta=False))
        t_2 = timeit(lambda: K._C.gemm(a, b, two_cta=True))
        t_cb = timeit(lambda: a @ b.t())
        row = {"name": name, "M": m, "N": n, "K": k, "tc_ms": round(t_tc, 4), "cublas_ms": round(t_cb, 4),
               "tc_tflops": round(fl / t_tc / 1e9, 1), "tc2_tflops": round(fl / t_2 / 1e9, 1), "cublas_tflops": round(fl / t_cb / 1e9, 1)}
        # backward flavours
        dy = torch.randn(m, n, device="cuda").bfloat16()
        t_dg = timeit(lambda: K._C.gemm(dy, b, b_mn=True, two_cta=False))
        t_dg2 = timeit(lambda: K._C.gemm(dy, b, b_mn=True, two_cta=True))
        t_wg2 = timeit(lambda: K._C.gemm(dy, a, a_mn=True, b_mn=True, two_cta=True))
        t_dg_cb = timeit(lambda: dy @ b)
        t_wg = timeit(lambda: K._C.gemm(dy, a, a_mn=True, b_mn=True, two_cta=False))
        t_wg_cb = timeit(lambda: dy.t() @ a)
        row.update({"dgrad_tc2_tflops": round(fl / t_dg2 / 1e9, 1), "wgrad_tc2_tflops": round(fl / t_wg2 / 1e9, 1)})
        row.update({"dgrad_tc_tflops": round(fl / t_dg / 1e9, 1), "dgrad_cublas_tflops": round(fl / t_dg_cb / 1e9, 1),
                    "wgrad_tc_tflops": round(fl / t_wg / 1e9, 1), "wgrad_cublas_tflops": round(fl / t_wg_cb / 1e9, 1)})
        rows.append(row)
        print(json.dumps(row), flush=True)
    return rows


if __name__ == "__main__":
    main()
