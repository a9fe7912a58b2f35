"""Launch each hot kernel a few times at the 1.3B shapes so ncu can capture them (see bench/capture_profiles.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from zero_transformer_b200.ops import cuda as K
from zero_transformer_b200.ops.attention_cuda import alibi_attention
from zero_transformer_b200.ops.reference import get_slopes
from zero_transformer_b200.parallel.fused import pack_segments

which = sys.argv[1] if len(sys.argv) > 1 else "all"
M, D = 16384, 2048
dev = "cuda"
rng = K.rng_state(dev)
if which in ("gemm", "all"):
    x = torch.randn(M, D, device=dev).bfloat16()
    w = torch.randn(3 * D, D, device=dev).bfloat16() * 0.02
    w1 = torch.randn(4 * D, D, device=dev).bfloat16() * 0.02
    res = torch.randn(M, D, device=dev).bfloat16()
    dy = torch.randn(M, 3 * D, device=dev).bfloat16()
    z = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)
    g = torch.zeros(3 * D, D, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        K._C.gemm(x, w)                                                    # fwd qkv
        K._C.gemm(x, w1, epilogue=K.EPI_GELU, aux_out=z)                   # fc_in + GELU
        K._C.gemm(z, w1.t().contiguous(), epilogue=K.EPI_DROPOUT_ADD, aux_in=res, p=0.1, rng=rng, op_id=1)  # fc_out + dropout + residual
        K._C.gemm(dy, w, b_mn=True)                                        # dgrad
        K._C.gemm(dy, x, C=g, accumulate=True, a_mn=True, b_mn=True)       # wgrad, accumulate
if which in ("attn", "all"):
    B, T, H, hd = 16, 1024, 16, 128
    qkv = torch.randn(B, T, 3 * H * hd, device=dev).bfloat16().requires_grad_(True)
    slopes = torch.tensor(get_slopes(H), device=dev)
    dout = torch.randn(B, T, H * hd, device=dev).bfloat16()
    for _ in range(3):
        K.begin_forward(dev)
        o = alibi_attention(qkv, slopes, H, hd ** -0.5, 0.1)
        o.backward(dout)
if which in ("misc", "all"):
    n = 1311082496 // 24  # one block bucket worth of parameters
    n -= n % 2048
    master = torch.randn(n, device=dev)
    m, v = torch.zeros_like(master), torch.zeros_like(master)
    fp = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    gr = torch.randn(n, device=dev).bfloat16()
    segs, tiles = pack_segments([(0, 0, n, 1)], dev)
    x = torch.randn(M, D, device=dev).bfloat16()
    wln = torch.ones(D, device=dev).bfloat16()
    for s in range(1, 4):
        K._C.adamw_shard(master, m, v, fp, [gr.data_ptr()], False, segs, 1, tiles, 1e-4, 0.9, 0.95, 1e-8, 0.1, 1.0, 1.0, s, 0, 0)
        y, mu, rs = K._C.layernorm_fwd(x, wln, 1e-5)
        K._C.layernorm_bwd(y, x, wln, mu, rs)
    logits = torch.randn(M, 50304, device=dev).bfloat16()
    labels = torch.randint(0, 50304, (M,), device=dev, dtype=torch.int32)
    for _ in range(2):
        lse, nll = K._C.ce_fwd(logits, labels)
        K._C.ce_bwd_(logits, labels, lse, torch.ones(1, device=dev), 1.0 / M)
torch.cuda.synchronize()
