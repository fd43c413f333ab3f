"""Raw NCCL all-reduce time of the flat fp32 gradient buffer (and of its post-backward remainder)
on the default communicator and on CTA-capped ones. torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from hero_b200 import distributed as hdist

rank, world, local = hdist.init()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)


def timeit(buf, group=None, n=10):
    for _ in range(3):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


full = torch.zeros(107_600_000, device=dev)
tail = full[:44_000_000]
groups = {"default": None}
for c in (2, 4, 8, 16):
    o = dist.ProcessGroupNCCL.Options()
    o.config.max_ctas = c
    o.config.min_ctas = 1
    groups[f"max_ctas={c}"] = dist.new_group(backend="nccl", pg_options=o)
for name, g in groups.items():
    a, b = timeit(full, g), timeit(tail, g)
    if rank == 0:
        print(f"{name:12s}: 430 MB {a:.3f} ms ({0.4304 / a:.0f} GB/s algbw)   176 MB {b:.3f} ms",
              flush=True)
dist.destroy_process_group()
