"""Data-parallel plumbing: the function names of the reference's utils/distributed.py on top of
torch.distributed (NCCL over NVLink 5 / NVSwitch on the B200 box, gloo in CPU tests) instead of
Horovod.

Semantics kept from the reference:
  * `all_reduce_and_rescale_tensors(tensors, denom)`: Horovod's `allreduce_` AVERAGES, so the
    result is mean-over-ranks / denom (utils/distributed.py:19-46).
  * `broadcast_tensors(tensors, root)` (utils/distributed.py:103-151), `all_gather_list`,
    `any_broadcast` (utils/distributed.py:182-212).
  * `VsmAllgather`: forward all-gather, backward = own slice of the incoming gradient with no
    reduction (model/pretrain.py:427-447).
What changes is the mechanism: when the gradients already live in one flat buffer
(`FlatParams.ensure_flat_grads`) the exchange is ONE in-place `all_reduce(AVG)` on that buffer —
no pack / unpack copies; otherwise tensors are coalesced once. One process per GPU, launched by
torchrun-style environment variables.
"""
import os
import pickle

import torch
import torch.distributed as dist


def init(backend=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).
    Returns (rank, world_size, local_rank); a no-op single-process job when unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def size():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def _avg_inplace(buf):
    if size() == 1:
        return
    if dist.get_backend() == "nccl":
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)     # ncclAvg: in-switch (NVLS) capable
    else:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        buf.div_(size())


def all_reduce_flat(buf, rescale_denom=1.0):
    """Mean over ranks of a flat gradient buffer, in place."""
    _avg_inplace(buf)
    if rescale_denom != 1.0:
        buf.div_(rescale_denom)
    return buf


def all_reduce_and_rescale_tensors(tensors, rescale_denom):
    """utils/distributed.py:19-46. Tensors that are consecutive views of one storage are reduced
    in place as a single message; anything else is coalesced once."""
    tensors = list(tensors)
    if not tensors:
        return
    flat = _as_single_view(tensors)
    if flat is not None:
        all_reduce_flat(flat, float(rescale_denom))
        return
    buf = torch.cat([t.reshape(-1) for t in tensors])
    all_reduce_flat(buf, float(rescale_denom))
    off = 0
    for t in tensors:
        n = t.numel()
        t.view(-1).copy_(buf[off:off + n])
        off += n


def _as_single_view(tensors):
    """If the tensors tile one contiguous region of a single storage (gaps allowed only as the
    alignment padding of FlatParams, which holds zeros), return that region as one tensor."""
    t0 = tensors[0]
    if not all(t.is_contiguous() and t.dtype == t0.dtype and t.device == t0.device and
               t.untyped_storage().data_ptr() == t0.untyped_storage().data_ptr()
               for t in tensors):
        return None
    es = t0.element_size()
    lo = min(t.data_ptr() for t in tensors)
    hi = max(t.data_ptr() + t.numel() * es for t in tensors)
    covered = sum(t.numel() for t in tensors)
    total = (hi - lo) // es
    if total > covered + 64 * len(tensors):     # more than alignment slack: not one region
        return None
    base = t0.untyped_storage().data_ptr()
    out = torch.empty(0, dtype=t0.dtype, device=t0.device)
    out.set_(t0.untyped_storage(), (lo - base) // es, (total,), (1,))
    return out


def broadcast_tensors(tensors, root_rank, buffer_size=None):
    """utils/distributed.py:103-151: every rank ends with root's values."""
    if size() == 1:
        return
    tensors = list(tensors)
    flat = _as_single_view(tensors) if tensors else None
    if flat is not None:
        dist.broadcast(flat, root_rank)
        return
    for t in tensors:
        dist.broadcast(t, root_rank)


def all_gather_list(data):
    """utils/distributed.py:182-198: gather arbitrary picklable data from all ranks. Stays off the
    GPU critical path (object collectives), unlike the reference's byte-tensor + .item() version."""
    if size() == 1:
        return [pickle.loads(pickle.dumps(data))]
    out = [None] * size()
    dist.all_gather_object(out, data)
    return out


def any_broadcast(data, root_rank):
    """utils/distributed.py:201-212."""
    if size() == 1:
        return pickle.loads(pickle.dumps(data))
    box = [data if rank() == root_rank else None]
    dist.broadcast_object_list(box, src=root_rank)
    return box[0]


class VsmAllgather(torch.autograd.Function):
    """model/pretrain.py:427-447: all-gather along dim 0 in rank order (ranks may contribute
    different row counts, as hvd.allgather allows); the backward hands each rank the slice of the
    gradient that corresponds to its own rows (no reduction)."""

    @staticmethod
    def forward(ctx, tensor, name=None):
        ctx.span = (0, tensor.shape[0])
        if size() == 1:
            return tensor
        tensor = tensor.contiguous()
        n = torch.tensor([tensor.shape[0]], dtype=torch.int64, device=tensor.device)
        counts = [torch.zeros_like(n) for _ in range(size())]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        rest = tuple(tensor.shape[1:])
        if len(set(counts)) == 1:
            out = torch.empty((sum(counts),) + rest, dtype=tensor.dtype, device=tensor.device)
            dist.all_gather_into_tensor(out, tensor)
        else:
            mx = max(counts)
            padded = tensor.new_zeros((mx,) + rest)
            padded[:tensor.shape[0]] = tensor
            parts = [torch.empty_like(padded) for _ in range(size())]
            dist.all_gather(parts, padded)
            out = torch.cat([p[:c] for p, c in zip(parts, counts)], 0)
        start = sum(counts[:rank()])
        ctx.span = (start, start + counts[rank()])
        return out

    @staticmethod
    def backward(ctx, grad_output):
        a, b = ctx.span
        return grad_output[a:b], None


vsm_allgather = VsmAllgather.apply
