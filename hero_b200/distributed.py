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


class GradBucketer:
    """Overlaps the data-parallel gradient exchange with the rest of the backward pass.

    The reference all-reduces one flat copy of every gradient AFTER backward has finished
    (utils/distributed.py:19-46, called at train_vcmr.py:233-239). Here gradients are accumulated
    in place in FlatParams' flat fp32 buffer, and the hand-written transformer backward reports
    each layer's parameters as soon as their gradient is final: the ranges of the flat buffer they
    cover are all-reduced (mean) right away on the communicator's stream while the layers below
    are still being differentiated. `finish()` reduces every range not covered so far and waits.
    The result equals one all_reduce_flat(grad_flat) after backward.

    Usage (one step):
        with bucketer:                 # installs the hook for forward + backward
            loss = model(...); loss.backward()
        bucketer.finish(); optimizer.step()

    A parameter used by several forward calls (e.g. c_encoder for video rows and again for query
    rows) is exchanged only after its LAST backward: forward registers every use (`expect`),
    backward retires them (`ready`)."""

    def __init__(self, flat, min_elems=1 << 20):
        self.flat = flat
        self.min_elems = min_elems       # merge announced ranges into messages of >= 4 MB
        self.reset()

    def reset(self):
        self.pending = {}       # id(param) -> forward uses whose backward has not run yet
        self.done = []          # [a, b) ranges of the flat buffer already handed to the backend
        self.handles = []
        self.queue = []         # final but not yet sent ranges (waiting to reach min_elems)

    # ---- hook protocol (hero_b200.functional.GRAD_HOOK) -------------------------------------
    def expect(self, params):
        for p in params:
            self.pending[id(p)] = self.pending.get(id(p), 0) + 1

    def ready(self, params):
        gf = self.flat.grad_flat
        if gf is None or size() == 1:
            return
        base, spans = gf.data_ptr(), []
        for p in params:
            k = id(p)
            left = self.pending.get(k, 1) - 1
            self.pending[k] = left
            ent = self.flat._by_id.get(k)
            if left > 0 or ent is None or p.grad is None:
                continue
            if p.grad.data_ptr() != base + 4 * ent[0]:
                continue                # gradient does not live in the flat buffer: finish() only
            spans.append(ent)
        for off, n in sorted(spans):
            end = min((off + n + 63) // 64 * 64, self.flat.total)   # alignment padding is zeros
            if self.queue and off <= self.queue[-1][1]:
                self.queue[-1][1] = max(self.queue[-1][1], end)
            else:
                self.queue.append([off, end])
        if sum(b - a for a, b in self.queue) >= self.min_elems:
            self._flush()

    def __enter__(self):
        from . import functional
        self.reset()
        functional.GRAD_HOOK[0] = self
        return self

    def __exit__(self, *exc):
        from . import functional
        functional.GRAD_HOOK[0] = None
        return False

    # ---- exchange ------------------------------------------------------------------------------
    def _launch(self, a, b):
        if b <= a:
            return
        buf = self.flat.grad_flat[a:b]
        if dist.get_backend() == "nccl":
            self.handles.append((dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True), None))
        else:
            self.handles.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True), buf))
        self.done.append((a, b))

    def _flush(self):
        for a, b in self.queue:
            self._launch(a, b)
        self.queue = []

    def finish(self, rescale_denom=1.0):
        """Exchange every range not sent so far, wait for all of it (the current stream waits; the
        host does not block on NCCL) and apply the reference's rescale."""
        if size() > 1:
            self.flat.ensure_flat_grads()
            self._flush()
            pos = 0
            for a, b in sorted(self.done):
                if a > pos:
                    self._launch(pos, a)
                pos = max(pos, b)
            if pos < self.flat.total:
                self._launch(pos, self.flat.total)
            for h, buf in self.handles:
                h.wait()
                if buf is not None:
                    buf.div_(size())
        if rescale_denom != 1.0:
            self.flat.grad_flat.div_(rescale_denom)
        self.reset()


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
