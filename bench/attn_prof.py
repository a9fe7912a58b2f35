import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zero_transformer_b200.ops import cuda as K
from zero_transformer_b200.ops.attention_cuda import alibi_attention
from zero_transformer_b200.ops.reference import get_slopes
B, T, H, hd = 16, 1024, 16, 128
qkv = torch.randn(B, T, 3 * H * hd, device="cuda").bfloat16().requires_grad_(True)
slopes = torch.tensor(get_slopes(H), device="cuda")
dout = torch.randn(B, T, H * hd, device="cuda").bfloat16()
for _ in range(2):
    K.begin_forward("cuda")
    o = alibi_attention(qkv, slopes, H, hd ** -0.5, 0.0)
    o.backward(dout)
torch.cuda.synchronize()
