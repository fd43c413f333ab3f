"""GradBucketer with the peer-copy transport == one NCCL all-reduce of the flat gradient buffer.
torchrun --nproc-per-node N tools/p2p_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from hero_b200 import distributed as hdist
from hero_b200.params import flat_of

rank, world, local = hdist.init()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
model = bench.build_model(dev)
flat = flat_of(model, dev)
bucketer = hdist.GradBucketer(flat, transport="p2p")
g = flat.ensure_flat_grads()
assert g.data_ptr() == bucketer.p2p.grad.data_ptr()
worst = 0.0
for trial in range(3):
    gen = torch.Generator(device=dev).manual_seed(100 * trial + rank)
    g.copy_(torch.randn(g.numel(), device=dev, generator=gen))
    ref = g.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    # uneven buckets (64-aligned), two holes left for finish()
    cuts = sorted({0, 64 * 1000, 64 * 50_000, 64 * 50_001, 64 * 400_000, 64 * 900_000,
                   64 * 1_300_000, flat.total})
    ranges = [(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    bucketer.reset()
    for i, (a, b) in enumerate(ranges):
        if i in (1, len(ranges) - 1):
            continue                      # left to finish()
        bucketer.queue.append([a, b])
        bucketer._flush()
    bucketer.finish()
    torch.cuda.synchronize()
    err = float((g - ref).abs().max())
    worst = max(worst, err)
    assert err <= 1e-6, (trial, err)
t = torch.tensor([worst], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"p2p_check ok: world={world} max |p2p - nccl| = {float(t):.2e}")
dist.destroy_process_group()
