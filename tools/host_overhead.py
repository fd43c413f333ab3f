"""Host enqueue time vs GPU time per bench step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hero_b200 import synth
from hero_b200.params import flat_of
from hero_b200.plan import attach_plan

dev = torch.device("cuda:0")
model = bench.build_model(dev)
flat = flat_of(model, dev)
gflat = flat.ensure_flat_grads()
vb, qb = synth.syn_tvr_dense(batch_size=32, seed=1234)
vbd = synth.to_device(attach_plan(dict(vb)), dev)
qbd = synth.to_device(attach_plan(dict(qb), kind="txt"), dev)
dclip = torch.randn(32, 100, 768, device=dev) * 1e-2
dq = torch.randn(32, 16, 768, device=dev) * 1e-2


def step():
    gflat.zero_()
    clip = model(vbd, "repr")
    q = model.f_encoder(qbd, "txt")[0]
    torch.autograd.backward([clip, q], [dclip, dq])


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {1e3 * t_host / n:.2f} ms/step; wall incl. GPU drain {1e3 * t_all / n:.2f} ms/step")
# forward only
with torch.no_grad():
    for _ in range(3):
        model(vbd, "repr")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        model(vbd, "repr"); model.f_encoder(qbd, "txt")
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    ta = time.perf_counter() - t0
print(f"fwd-only (no_grad, train-mode dropout): host {1e3 * th / n:.2f} ms, wall {1e3 * ta / n:.2f} ms")
