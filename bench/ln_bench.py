"""LayerNorm-backward variants at the 1.3B shape (M = 16384 rows, D = 2048), CUDA-event timed, inputs produced by a
preceding kernel the way the training step does (so part of the traffic is L2-resident, as in the step)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from zero_transformer_b200.ops import cuda as K

M, D = 16384, 2048
x = torch.randn(M, D, device="cuda").bfloat16()
dy = torch.randn(M, D, device="cuda").bfloat16()
dres = torch.randn(M, D, device="cuda").bfloat16()
w = torch.ones(D, device="cuda").bfloat16()
y, mean, rstd = K._C.layernorm_fwd(x, w, 1e-5)
g = torch.zeros(D, device="cuda", dtype=torch.bfloat16)
scratch = torch.zeros(D, device="cuda")
ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
rng = K.rng_state("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20, cold=False):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


for cold in (False, True):
    r = {"cold_l2": cold}
    r["fwd_us"] = round(timeit(lambda: K._C.layernorm_fwd(x, w, 1e-5), cold=cold), 1)
    r["bwd_plain_us"] = round(timeit(lambda: K._C.layernorm_bwd(dy, x, w, mean, rstd, dres), cold=cold), 1)
    r["bwd_grad_in_kernel_us"] = round(timeit(lambda: K._C.layernorm_bwd(dy, x, w, mean, rstd, dres, grad_out=g, accumulate=True, scratch=scratch, ticket=ticket), cold=cold), 1)
    r["bwd_grad_plus_dropout_mask_us"] = round(timeit(lambda: K._C.layernorm_bwd(dy, x, w, mean, rstd, dres, grad_out=g, accumulate=True, scratch=scratch, ticket=ticket,
                                                                                     drop_p=0.1, rng=rng, op_id=3), cold=cold), 1)
    r["standalone_dropout_bwd_us"] = round(timeit(lambda: K._C.dropout_add(dy, None, 0.1, rng, 3), cold=cold), 1)
    gb = 4 * M * D * 2 / 1e9
    r["bwd_plain_gbs"] = round(gb / (r["bwd_plain_us"] * 1e-6), 0)
    print(json.dumps(r), flush=True)
