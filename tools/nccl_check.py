"""Run under torchrun on N GPUs: checks the NCCL data-parallel plumbing of hero_b200.distributed
(mean all-reduce of the flat gradient buffer, broadcast, VsmAllgather) and that every rank ends an
optimizer step with identical parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/nccl_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from hero_b200 import distributed as hd
from hero_b200 import synth
from hero_b200.optim import FusedAdamW
from hero_b200.params import flat_of


def main():
    rank, world, local = hd.init()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # 1. mean all-reduce == manual mean of rank-dependent buffers (fp32: exact for these values)
    buf = torch.arange(1 << 20, dtype=torch.float32, device=dev) * (rank + 1)
    hd.all_reduce_flat(buf)
    expect = torch.arange(1 << 20, dtype=torch.float32, device=dev) * (sum(range(1, world + 1)) / world)
    assert torch.allclose(buf, expect, rtol=1e-6), "mean all-reduce mismatch"
    # 2. VsmAllgather forward/backward
    x = (torch.ones(3, 4, device=dev) * (rank + 1)).requires_grad_(True)
    y = hd.vsm_allgather(x)
    assert y.shape[0] == 3 * world and float(y[3 * rank, 0]) == rank + 1
    (y * torch.arange(y.numel(), device=dev).view_as(y)).sum().backward()
    assert torch.equal(x.grad, torch.arange(y.numel(), device=dev).view_as(y)[3 * rank:3 * rank + 3].float())
    # 3. a data-parallel training step: different data per rank, identical params afterwards
    model = bench.build_model(dev, seed=rank)          # deliberately different init per rank
    flat = flat_of(model, dev)
    hd.broadcast_tensors([flat.flat], 0)
    flat.mark_dirty()
    gflat = flat.ensure_flat_grads()
    opt = FusedAdamW(flat, lr=1e-3)
    vb, qb = synth.syn_tvr_ragged(batch_size=4, seed=100 + rank, t_range=(20, 40), s_range=(4, 8),
                                  l_range=(4, 20))
    clip, q = model.forward_repr_txt(synth.to_device(vb, dev), synth.to_device(qb, dev))
    (clip.float().pow(2).mean() + q.float().pow(2).mean()).backward()
    hd.all_reduce_flat(gflat)
    gn = opt.clip_grad_norm_(1.0)
    opt.step()
    digest = torch.stack([flat.flat.double().sum(), flat.flat.double().abs().sum(),
                          gflat.double().abs().sum()])
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    for g in gathered:
        assert torch.equal(g, gathered[0]), "ranks diverged after the data-parallel step"
    if rank == 0:
        print(f"nccl_check ok: world={world} grad_norm={gn:.4f} param_digest={gathered[0].tolist()}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
