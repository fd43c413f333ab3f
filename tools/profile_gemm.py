"""A handful of encoder-shaped GEMM launches for `ncu --set full` (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hero_b200 import ops

dev = torch.device("cuda:0")
M = 16000
x = torch.randn(M, 768, device=dev).bfloat16()
w1 = (torch.randn(3072, 768, device=dev) * 0.05).bfloat16()
b1 = torch.randn(3072, device=dev)
wq = (torch.randn(2304, 768, device=dev) * 0.05).bfloat16()
bq = torch.randn(2304, device=dev)
f = torch.empty(M, 3072, dtype=torch.bfloat16, device=dev)
pre = torch.empty_like(f)
qkv = torch.empty(M, 2304, dtype=torch.bfloat16, device=dev)
dw = torch.zeros(3072, 768, device=dev)
wo = (torch.randn(768, 768, device=dev) * 0.05).bfloat16()
bo = torch.randn(768, device=dev)
h = torch.randn(M, 768, device=dev).bfloat16()
s1 = torch.empty_like(h)
drop = ops.drop_params(0.1, 99)
for _ in range(3):
    ops.gemm(x, wo, s1, bias=bo, resid=h, drop=drop)             # out-proj + dropout + residual
    ops.gemm(x, w1, f, bias=b1, act=ops.ACT_GELU, aux_out=pre)   # FFN-up + erf-GELU (training)
    ops.gemm(x, wq, qkv, bias=bq)                                 # fused QKV
    ops.gemm(f, w1, x, b_mn=True)                                 # dgrad-shaped (K = 3072)
    ops.gemm(f, x, dw, a_mn=True, b_mn=True, accumulate_f32=True) # wgrad (K = tokens)
torch.cuda.synchronize()
print("done")
