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
    # fused-epilogue variants at the fc shapes
    m, d = M, 2048
    x = torch.randn(m, d, device="cuda").bfloat16()
    w1 = torch.randn(4 * d, d, device="cuda").bfloat16() * 0.02
    w2 = torch.randn(d, 4 * d, device="cuda").bfloat16() * 0.02
    z = torch.randn(m, 4 * d, device="cuda").bfloat16()
    dy = torch.randn(m, d, device="cuda").bfloat16()
    res = torch.randn(m, d, device="cuda").bfloat16()
    rng = K.rng_state("cuda")
    fl = 2.0 * m * d * 4 * d
    for two in (False, True):
        r = {"two_cta": two}
        r["fc_in_gelu"] = round(fl / timeit(lambda: K._C.gemm(x, w1, epilogue=K.EPI_GELU, aux_out=z, two_cta=two)) / 1e9, 1)
        r["fc_out_dropout_add"] = round(fl / timeit(lambda: K._C.gemm(z, w2, epilogue=K.EPI_DROPOUT_ADD, aux_in=res, p=0.1, rng=rng, op_id=1, two_cta=two)) / 1e9, 1)
        r["fc_out_plain"] = round(fl / timeit(lambda: K._C.gemm(z, w2, two_cta=two)) / 1e9, 1)
        r["dgrad_gelu_bwd"] = round(fl / timeit(lambda: K._C.gemm(dy, w2, epilogue=K.EPI_GELU_BWD, aux_in=z, b_mn=True, two_cta=two)) / 1e9, 1)
        r["dgrad_plain"] = round(fl / timeit(lambda: K._C.gemm(dy, w2, b_mn=True, two_cta=two)) / 1e9, 1)
        print(json.dumps(r), flush=True)
    # ---- round 2: staging-free (direct st.global) outputs and the vocabulary raster
    for direct in (0, 1, 2, 3):
        r = {"direct_store": direct}
        r["fc_in_gelu"] = round(fl / timeit(lambda: K._C.gemm(x, w1, epilogue=K.EPI_GELU, aux_out=z, two_cta=True, direct_store=direct)) / 1e9, 1)
        r["fc_out_dropout_add"] = round(fl / timeit(lambda: K._C.gemm(z, w2, epilogue=K.EPI_DROPOUT_ADD, aux_in=res, p=0.1, rng=rng, op_id=1, two_cta=True, direct_store=direct)) / 1e9, 1)
        r["dgrad_gelu_bwd"] = round(fl / timeit(lambda: K._C.gemm(dy, w2, epilogue=K.EPI_GELU_BWD, aux_in=z, b_mn=True, two_cta=True, direct_store=direct)) / 1e9, 1)
        r["fc_in_plain"] = round(fl / timeit(lambda: K._C.gemm(x, w1, two_cta=True, direct_store=direct)) / 1e9, 1)
        print(json.dumps(r), flush=True)
    V = 50304
    h = torch.randn(m, d, device="cuda").bfloat16()
    wte = torch.randn(V, d, device="cuda").bfloat16() * 0.02
    dl = torch.randn(m, V, device="cuda").bfloat16()
    flv = 2.0 * m * V * d
    for label, kw in (("1cta", dict(two_cta=False)), ("2cta_raster_m", dict(two_cta=True, raster_n=0)), ("2cta_raster_n", dict(two_cta=True, raster_n=1))):
        r = {"lm_head": label}
        r["fwd"] = round(flv / timeit(lambda: K._C.gemm(h, wte, **kw)) / 1e9, 1)
        r["dgrad"] = round(flv / timeit(lambda: K._C.gemm(dl, wte, b_mn=True, **kw)) / 1e9, 1)
        r["wgrad"] = round(flv / timeit(lambda: K._C.gemm(dl, h, a_mn=True, b_mn=True, **kw)) / 1e9, 1)
        print(json.dumps(r), flush=True)
    return rows


if __name__ == "__main__":
    main()
