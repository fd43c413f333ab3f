"""FFN-up + GELU and GELU'-multiply dgrad launched as CTA pairs, for ncu (why are pairs slower
there?). See profiles/README.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hero_b200 import ops

dev = torch.device("cuda:0")
M = 16512
x = torch.randn(M, 768, device=dev).bfloat16()
w1 = (torch.randn(3072, 768, device=dev) * 0.05).bfloat16()
b1 = torch.randn(3072, device=dev)
f = torch.empty(M, 3072, dtype=torch.bfloat16, device=dev)
pre = torch.empty_like(f)
g = torch.randn(M, 768, device=dev).bfloat16()
w2 = (torch.randn(768, 3072, device=dev) * 0.05).bfloat16()
dpre = torch.empty_like(f)
for _ in range(3):
    ops.gemm(x, w1, f, bias=b1, act=ops.ACT_GELU, aux_out=pre, cta_pair=2)
    ops.gemm(g, w2, dpre, b_mn=True, act=3, aux_in=pre, cta_pair=2)
    ops.gemm(x, w1, f, bias=b1, act=ops.ACT_GELU, aux_out=pre, cta_pair=1)
torch.cuda.synchronize()
print("done")
