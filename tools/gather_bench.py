"""gather_sum_rows (fp32 accumulate form: position-table gradients) at bench-like shapes:
60 table rows, one entry per sequence each (660 sequences), 13.3k x 768 bf16 source rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hero_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n_seq, L, H = 660, 20, 768
T = n_seq * L
src = torch.randn(T, H, device=dev).bfloat16()
pos = np.tile(np.arange(L), n_seq)                      # token -> table row
order = np.argsort(pos, kind="stable").astype(np.int32)
off = np.concatenate([[0], np.cumsum(np.bincount(pos, minlength=L))]).astype(np.int32)
off_d, idx_d = torch.from_numpy(off).to(dev), torch.from_numpy(order).to(dev)
dst = torch.zeros(L, H, device=dev)
for _ in range(5):
    ops.gather_sum_rows(src, off_d, idx_d, dst)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.gather_sum_rows(src, off_d, idx_d, dst)
e1.record()
torch.cuda.synchronize()
ref = torch.zeros(L, H, device=dev)
ref.index_add_(0, torch.from_numpy(pos).to(dev), src.float())
dst.zero_()
ops.gather_sum_rows(src, off_d, idx_d, dst)
print(f"gather_sum_rows f32: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call "
      f"({T * H * 2 / 1e6:.1f} MB read), max err {(dst - ref).abs().max().item():.3e}")
