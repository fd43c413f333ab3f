"""cProfile of the host side of a bench step (where do the enqueue milliseconds go?)."""
import cProfile, io, os, pstats, sys
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
    clip, q = model.forward_repr_txt(vbd, qbd)
    torch.autograd.backward([clip, q], [dclip, dq])


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
