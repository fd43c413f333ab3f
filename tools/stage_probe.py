"""Where does BatchStager.stage spend host time? (pinned copies vs plan index upload)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hero_b200 import synth
from hero_b200.loader import BatchStager
from hero_b200.plan import DeviceIndex, build_plans, plan_inputs

dev = torch.device("cuda:0")
vb, qb = synth.syn_tvr_dense(batch_size=32, seed=1)
for b in (vb, qb):
    for k, v in b.items():
        if torch.is_tensor(v):
            b[k] = v.pin_memory()
st = BatchStager(dev, 3)


def t(fn, n=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return dt


def only_tensors():
    _, _, slot = st.stage(vb, qb)
    st.release(slot)


print(f"stage tensors only      : {t(only_tensors):.3f} ms")
rp, tp = build_plans(plan_inputs(vb), plan_inputs(qb, 'txt'))
arrs = rp.f.arrays("f_")


def one_index():
    DeviceIndex(arrs, dev)


print(f"one DeviceIndex (f_)    : {t(one_index):.3f} ms  ({sum(a.size for a in arrs.values()) * 4 / 1e6:.2f} MB)")


def plans_to():
    for p in (rp, tp, rp.__dict__['_joint']):
        p.dev = None
        p.to(dev)


print(f"3 plan uploads          : {t(plans_to):.3f} ms")
for k, v in vb.items():
    if torch.is_tensor(v):
        buf = torch.empty_like(v, device=dev)
        print(f"  copy_ {k:22s} {tuple(v.shape)} {v.dtype} pinned={v.is_pinned()} "
              f"{t(lambda: buf.copy_(v, non_blocking=True)):.3f} ms")
