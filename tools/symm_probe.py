"""Viability probe: torch symmetric memory + peer copies on this box (torchrun, N ranks)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from hero_b200 import distributed as hdist

rank, world, local = hdist.init()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
g = dist.group.WORLD
try:
    symm.enable_symm_mem_for_group(g.group_name)
except Exception as e:  # noqa
    print("enable:", repr(e)[:200], flush=True)
n = 50_000_000
buf = symm.empty(n, dtype=torch.float32, device=dev)
h = symm.rendezvous(buf, g.group_name)
if rank == 0:
    print("rendezvous ok; multicast:", h.has_multicast_support, "signal pad", h.signal_pad_size,
          flush=True)
buf.fill_(float(rank + 1))
src = torch.full((n,), float(10 + rank), device=dev)
h.barrier(0)
peer = (rank + 1) % world
pbuf = h.get_buffer(peer, (n,), torch.float32, 0)
print(rank, "peer tensor device", pbuf.device, flush=True)
cudart = ctypes.CDLL("libcudart.so.12")


def t(fn, it=5):
    fn()
    torch.cuda.synchronize()
    h.barrier(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


ms = t(lambda: pbuf.copy_(src, non_blocking=True))
if rank == 0:
    print(f"torch copy_ to peer      : {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s", flush=True)
st = torch.cuda.current_stream().cuda_stream


def memcpy():
    rc = cudart.cudaMemcpyAsync(ctypes.c_void_p(pbuf.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                ctypes.c_size_t(n * 4), 3, ctypes.c_void_p(st))
    assert rc == 0, rc


ms = t(memcpy)
if rank == 0:
    print(f"cudaMemcpyAsync to peer  : {ms:.3f} ms  {n * 4 / ms / 1e6:.0f} GB/s", flush=True)
h.barrier(0)
torch.cuda.synchronize()
ok = bool((buf == float(10 + (rank - 1) % world)).all())
print(rank, "peer write landed:", ok, flush=True)
t0 = time.perf_counter()
for _ in range(20):
    h.put_signal(peer, 1)
    h.wait_signal((rank - 1) % world, 1)
torch.cuda.synchronize()
if rank == 0:
    print(f"put+wait signal round    : {(time.perf_counter() - t0) / 20 * 1e6:.1f} us", flush=True)
dist.destroy_process_group()
