"""Per-kernel device time of real (warm, back-to-back) bench steps through torch.profiler / CUPTI
activity tracing — unlike the ncu launch list the kernels are neither serialised nor cold-cache,
so the numbers add up to the measured step. Usage: python tools/step_profile.py [steps]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from hero_b200 import synth
from hero_b200.params import flat_of
from hero_b200.plan import attach_plan

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
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
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = ev.name.split("(")[0][:70]
    a = agg[name]
    a[0] += 1
    a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tr = ev.time_range
    t_min = tr.start if t_min is None else min(t_min, tr.start)
    t_max = tr.end if t_max is None else max(t_max, tr.end)
total = sum(a[1] for a in agg.values())
span = (t_max - t_min) / steps
print(f"| kernel | launches/step | us/step | avg us | share of busy |\n|---|---|---|---|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n / steps:.1f} | {us / steps:.1f} | {us / n:.1f} | {100 * us / total:.1f} % |")
print(f"\nbusy {total / steps:.0f} us/step, wall span {span:.0f} us/step over {steps} steps")

# timeline of the LAST step: kernel start (us since the step's first kernel), duration, gap to the
# previous kernel's end on the device (idle time if positive and no other stream is busy)
evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA),
             key=lambda ev: ev.time_range.start)
per = len(evs) // steps
last = evs[-per:]
t0 = last[0].time_range.start
out = os.environ.get("STEP_TIMELINE", "gpurun_out/step_timeline.csv")
busy_until, idle = t0, 0.0
with open(out, "w") as f:
    f.write("start_us,dur_us,idle_before_us,name\n")
    for ev in last:
        s, e = ev.time_range.start, ev.time_range.end
        gap = max(0.0, s - busy_until)
        idle += gap
        busy_until = max(busy_until, e)
        f.write(f"{s - t0:.1f},{e - s:.1f},{gap:.1f},{ev.name.split('(')[0][:80]}\n")
print(f"last step: {len(last)} kernels, device idle {idle:.0f} us of {busy_until - t0:.0f} us")
