"""Attention forward / backward at the bench shapes, back-to-back launches timed with CUDA events
(f-encoder rows: 25-token sequences, 16.5k tokens; temporal rows: 100-token sequences, 3.2k)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hero_b200 import ops  # noqa: E402
from hero_b200.plan import DeviceIndex, SeqPlan  # noqa: E402

dev = torch.device("cuda:0")
heads, H = 12, 768


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, lens in (("f rows 25 x 660", [25] * 660), ("c rows 100 x 32", [100] * 32),
                   ("ragged 5..70", list(np.random.RandomState(0).randint(5, 71, 420)))):
    mask = np.zeros((len(lens), max(lens)), np.int64)
    for r, n in enumerate(lens):
        mask[r, :n] = 1
    sp = SeqPlan(mask)
    att = sp.attn(DeviceIndex(sp.arrays("s_"), dev), "s_")
    M = sp.n_tok
    qkv = torch.randn(M, 3 * H, device=dev).bfloat16()
    ctx = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(M, heads, device=dev)
    dctx = torch.randn(M, H, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros(3 * H, device=dev)
    drop = ops.drop_params(0.1, 77)
    ops.attn_fwd(qkv, att, ctx, heads=heads, drop=drop, lse=lse)
    f_d = timed(lambda: ops.attn_fwd(qkv, att, ctx, heads=heads, drop=drop, lse=lse))
    f_e = timed(lambda: ops.attn_fwd(qkv, att, ctx, heads=heads))
    b_d = timed(lambda: ops.attn_bwd(qkv, att, ctx, dctx, lse, dqkv, heads=heads, drop=drop))
    b_db = timed(lambda: ops.attn_bwd(qkv, att, ctx, dctx, lse, dqkv, heads=heads, drop=drop,
                                      dbias=dbias))
    fwd_b, bwd_b = M * H * 2 * 4, M * H * 2 * 9       # qkv + ctx ; qkv + ctx + dctx + dqkv (+lse)
    print(f"{name}: M={M} tiles={sp.n_tiles}  fwd drop {f_d:.1f} us ({fwd_b / f_d / 1e6:.2f} TB/s)  "
          f"fwd eval {f_e:.1f} us  bwd {b_d:.1f} us ({bwd_b / b_d / 1e6:.2f} TB/s)  "
          f"bwd + bias grad {b_db:.1f} us")
