"""Event-timed encoder-shaped GEMMs, each in a loop of its own (no profiler): TF/s per shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hero_b200 import ops

dev = torch.device("cuda:0")


def bf(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).bfloat16()


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for M in (16512, 3200):
    x = bf(M, 768)
    w1, b1 = bf(3072, 768, s=0.05), torch.randn(3072, device=dev)
    wq, bq = bf(2304, 768, s=0.05), torch.randn(2304, device=dev)
    wo, bo = bf(768, 768, s=0.05), torch.randn(768, device=dev)
    w2, b2 = bf(768, 3072, s=0.05), torch.randn(768, device=dev)
    f, pre = torch.empty(M, 3072, dtype=torch.bfloat16, device=dev), torch.empty(M, 3072, dtype=torch.bfloat16, device=dev)
    qkv = torch.empty(M, 2304, dtype=torch.bfloat16, device=dev)
    h32 = torch.randn(M, 768, device=dev)
    s1 = torch.empty(M, 768, device=dev)
    s1b = torch.empty(M, 768, dtype=torch.bfloat16, device=dev)
    dw = torch.zeros(3072, 768, device=dev)
    dwq = torch.zeros(2304, 768, device=dev)
    dwo = torch.zeros(768, 768, device=dev)
    dx = torch.empty(M, 768, dtype=torch.bfloat16, device=dev)
    drop = ops.drop_params(0.1, 99)
    cases = [
        ("qkv  bias            ", 2 * M * 2304 * 768, lambda: ops.gemm(x, wq, qkv, bias=bq)),
        ("outp f32res drop     ", 2 * M * 768 * 768, lambda: ops.gemm(x, wo, s1, bias=bo, resid=h32, drop=drop)),
        ("outp bf16 plain      ", 2 * M * 768 * 768, lambda: ops.gemm(x, wo, s1b, bias=bo)),
        ("up   gelu+grad       ", 2 * M * 3072 * 768, lambda: ops.gemm(x, w1, f, bias=b1, act=ops.ACT_GELU, aux_out=pre)),
        ("up   gelu (infer)    ", 2 * M * 3072 * 768, lambda: ops.gemm(x, w1, f, bias=b1, act=ops.ACT_GELU)),
        ("down f32res drop     ", 2 * M * 768 * 3072, lambda: ops.gemm(f, w2, s1, bias=b2, resid=h32, drop=drop)),
        ("dgrad*gelu' N=3072   ", 2 * M * 3072 * 768, lambda: ops.gemm(x, w2, f, b_mn=True, act=ops.ACT_GELU_GRAD, aux_in=pre)),
        ("dgrad K=3072 +res    ", 2 * M * 768 * 3072, lambda: ops.gemm(f, w1, dx, b_mn=True, resid=x)),
        ("dgrad K=768          ", 2 * M * 768 * 768, lambda: ops.gemm(x, wo, dx, b_mn=True)),
        ("dgrad K=2304 +res    ", 2 * M * 768 * 2304, lambda: ops.gemm(qkv, wq, dx, b_mn=True, resid=x)),
        ("wgrad 3072x768       ", 2 * M * 3072 * 768, lambda: ops.gemm(f, x, dw, a_mn=True, b_mn=True, accumulate_f32=True)),
        ("wgrad 2304x768       ", 2 * M * 2304 * 768, lambda: ops.gemm(qkv, x, dwq, a_mn=True, b_mn=True, accumulate_f32=True)),
        ("wgrad 768x768        ", 2 * M * 768 * 768, lambda: ops.gemm(x, x, dwo, a_mn=True, b_mn=True, accumulate_f32=True)),
    ]
    for name, fl, fn in cases:
        us = timeit(fn)
        print(f"M={M:6d} {name} {us:8.1f} us  {fl / us / 1e6:7.0f} TF/s", flush=True)
