"""N = 1 cost of the data-parallel backward schedule (one native call per layer + hook) against the
whole-stack call — what the GradBucketer pays in compute before any communication happens."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hero_b200 import functional, synth
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


class NoopHook:
    def expect(self, params):
        pass

    def ready(self, params):
        pass


def step():
    gflat.zero_()
    clip, q = model.forward_repr_txt(vbd, qbd)
    torch.autograd.backward([clip, q], [dclip, dq])


def timeit(n=10):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"whole-stack backward : {timeit():.3f} ms/step")
functional.GRAD_HOOK[0] = NoopHook()
print(f"per-layer + hook     : {timeit():.3f} ms/step")
functional.GRAD_HOOK[0] = None
print(f"whole-stack backward : {timeit():.3f} ms/step")
