"""Headline benchmark: clips/sec, forward + backward of the HERO hierarchical encoder at
train-tvr-8gpu shapes (BASELINE.json), data-parallel over N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference          # CPU oracle port of the reference path (rank 0)

One step = one pass of the hot path over one synthetic SYN-TVR-dense batch per rank
(B = 32 clips x 100 frames x 4352-d features, 20 subtitle rows of 5 frames + 20 tokens per clip,
one 16-token query per clip; hero_finetune dims: 6 cross-modal + 3 temporal layers, H = 768):
HierarchicalVlModel 'repr' forward + CrossModalTrm 'txt' forward on the query rows (by default
through `forward_repr_txt`, which runs the query rows in the same cross-modal pass as the video
rows; `--separate-txt` issues the reference's two calls), backward of both from fixed upstream
gradients, and (N > 1) the mean all-reduce of the flat gradient buffer.
Training mode (dropout 0.1 as in config/train-tvr-8gpu.json). No optimizer step (the metric is
fwd+bwd); `--with-optimizer` adds the fused AdamW.

`value`   whole-job clips/s with inputs (and the per-batch packing plan, a collate-side product)
          resident in HBM, CUDA-event timed, max over ranks.
`e2e`     same metric through the public module API from PINNED HOST batches: per step the packing
          plan is built from that step's masks in loader worker processes (hero_b200.plan.PlanPool;
          the reference builds its gather indices in DataLoader collate workers), the batch and the
          plan's index arrays are copied host->device on a side stream (prefetch one step ahead,
          like the reference's PrefetchLoader, data/loader.py:89-144) and a loss scalar is read
          back. The loader is primed (two plans in flight) when the timed region starts.
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, INTER, HEADS, F_LAYERS, C_LAYERS, D = 768, 3072, 12, 6, 3, 4352
METRIC = "clips/sec fwd+bwd HERO encoder (TVR 8gpu shapes)"


def model_json(path):
    def cfg(n, with_vocab):
        c = {"attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1,
             "hidden_size": H, "initializer_range": 0.02, "intermediate_size": INTER,
             "max_position_embeddings": 514, "num_attention_heads": HEADS, "num_hidden_layers": n,
             "type_vocab_size": 2}
        if with_vocab:
            c["vocab_size"] = 50272
        return c
    with open(path, "w") as f:
        json.dump({"f_config": cfg(F_LAYERS, True), "c_config": cfg(C_LAYERS, False)}, f)


def algorithmic_flops_fwd(vb, qb):
    """SURVEY.md §8d counting rule: valid tokens only, GEMMs + QK^T + PV, multiply-add = 2."""
    def f_layer(n):
        return 24 * n * H * H + 4 * n * n * H
    f_lens = vb["f_attn_masks"].sum(1).tolist()
    c_lens = vb["c_attn_masks"].sum(1).tolist()
    q_lens = qb["attn_masks"].sum(1).tolist()
    n_img = sum(len(fr) for clip in vb["sub_idx2frame_idx"] for _, fr in clip)
    fl = F_LAYERS * sum(f_layer(n) for n in f_lens) + F_LAYERS * sum(f_layer(n) for n in q_lens)
    fl += C_LAYERS * sum(f_layer(n) for n in c_lens)
    fl += 2 * D * H * (n_img + sum(c_lens))
    return float(fl)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_model(device, seed=0):
    from hero_b200.model import HierarchicalVlModel, VideoModelConfig
    torch.manual_seed(seed)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "hero_finetune.json")
        model_json(p)
        model = HierarchicalVlModel(VideoModelConfig(p), vfeat_dim=D, max_frm_seq_len=100)
    model.initialize()   # random init of the hero_finetune architecture (no checkpoints offline)
    return model.to(device).train()


def run_ours(args):
    from hero_b200 import distributed as hdist
    from hero_b200 import ops, synth
    from hero_b200.params import flat_of
    from hero_b200.loader import BatchStager, record_plans
    from hero_b200.plan import PlanPool, attach_plan
    from hero_b200.optim import FusedAdamW

    rank, world, local_rank = hdist.init()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    model = build_model(device, seed=0)
    flat = flat_of(model, device)
    hdist.broadcast_tensors([flat.flat], 0)
    flat.mark_dirty()
    # N > 1: gradient buckets travel during backward (peer copies over NVLink; see GradBucketer);
    # built first because it moves the flat gradient buffer into symmetric memory
    bucketer = None
    if world > 1 and not args.no_overlap and not args.dp_skip_exchange:
        bucketer = hdist.overlapped_exchange(flat, transport=args.dp_transport,
                                             min_elems=args.bucket_elems,
                                             overlap_ctas=args.overlap_ctas)
    if world > 1 and args.dp_skip_exchange and os.environ.get("HERO_DP_DIAG") == "symm":
        hdist.PeerExchange(flat)     # diagnostic: gradients in symmetric memory, no exchange
    gflat = flat.ensure_flat_grads()
    opt = FusedAdamW(flat, lr=1e-4) if args.with_optimizer else None

    B = args.batch_size
    n_host = 3
    host = []
    for i in range(n_host):
        vb, qb = synth.syn_tvr_dense(batch_size=B, seed=1234 + rank + 1000 * i)
        for b in (vb, qb):
            for k, v in b.items():
                if torch.is_tensor(v):
                    b[k] = v.pin_memory()
        host.append((vb, qb))
    flops_fwd = algorithmic_flops_fwd(*host[0])
    g = torch.Generator().manual_seed(7)
    dclip = (torch.randn(B, 100, H, generator=g) * 1e-2).to(device)
    dq = (torch.randn(B, host[0][1]["input_ids"].shape[1], H, generator=g) * 1e-2).to(device)

    def fwd_bwd(vb_dev, qb_dev):
        if bucketer is not None:   # per-layer gradient exchange overlapped with backward
            bucketer.__enter__()
        if args.separate_txt:      # the reference's two calls (model/pretrain.py:65-70)
            clip = model(vb_dev, "repr")
            q = model.f_encoder(qb_dev, "txt")[0]
        else:                      # same results, query rows share the video rows' GEMMs
            clip, q = model.forward_repr_txt(vb_dev, qb_dev)
        torch.autograd.backward([clip, q], [dclip, dq])
        if bucketer is not None:
            bucketer.__exit__(None, None, None)
            bucketer.finish()
        elif world > 1 and not args.dp_skip_exchange:
            hdist.all_reduce_flat(gflat)
        if opt is not None:
            opt.step()
        return clip

    # ------------------------------------------------------------- device-resident timing
    resident = []
    for vb, qb in host:
        vb = attach_plan(dict(vb))
        qb = attach_plan(dict(qb), kind="txt")
        resident.append((synth.to_device(vb, device), synth.to_device(qb, device)))
    torch.cuda.synchronize()

    def resident_step(i):
        gflat.zero_()
        vb_dev, qb_dev = resident[i % n_host]
        fwd_bwd(vb_dev, qb_dev)

    for i in range(args.warmup):
        resident_step(i)
    ops.reset_launch_count()
    sampler = ClockSampler(local_rank)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        resident_step(i)
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / max(args.steps, 1)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = ops.launch_count() // max(args.steps, 1)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ------------------------------------------------------------- GEMM-family roofline (live)
    # (the layer runtime keeps everything on one stream while GEMM launches are being timed, so
    # durations do not overlap; the share below is taken against THIS pass's own step time)
    ops.start_gemm_profile()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(min(args.steps, 5)):
        resident_step(i)
    p1.record()
    torch.cuda.synchronize()
    gp = ops.stop_gemm_profile()
    profiled_step_ms = p0.elapsed_time(p1) / max(min(args.steps, 5), 1)
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = gp["flops"] / (gp["ms"] * 1e-3) / 1e12 if gp["ms"] > 0 else 0.0
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all variants)",
                "achieved": round(achieved, 1), "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(achieved / peak_tf, 4),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
                if peaks else "fallback 1.4 PF/s (of fallback)",
                "launches_per_step": gp["launches"] // max(min(args.steps, 5), 1),
                "gemm_share_of_step": round(gp["ms"] / max(min(args.steps, 5), 1) /
                                            profiled_step_ms, 3),
                "traffic": traffic,
                "step_algorithmic_tflops": round(3 * flops_fwd * 1e-12, 4),
                "step_frac_of_peak": round(3 * flops_fwd / (ms_per_step * 1e-3) / 1e12 / peak_tf,
                                           4)}

    # ------------------------------------------------------------- end-to-end from pinned host
    stager = BatchStager(device, depth=3)

    def h2d_bytes(b):
        return sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v))

    # Packing plans are built per step, from that step's masks, in worker processes — the
    # reference builds its gather indices in DataLoader collate workers (data/data.py) — two steps
    # ahead of their use; the training process uploads the finished index arrays.
    # (If the worker pool cannot start or dies, plans are built in this process instead — slower,
    # but the run still produces its line; `config.e2e_plans` says which.)
    loader_state = {"pool": None, "mode": "loader worker processes (PlanPool, 3 workers)"}
    try:
        loader_state["pool"] = PlanPool(workers=3)
    except Exception as e:                                  # noqa: BLE001
        loader_state["mode"] = f"in-process (worker pool unavailable: {type(e).__name__})"
    plan_futs = collections.deque()

    def submit_plan(i):
        vb, qb = host[i % n_host]
        pool = loader_state["pool"]
        try:
            plan_futs.append(pool.submit(vb, qb) if pool is not None else None)
        except Exception as e:                              # noqa: BLE001
            loader_state["pool"] = None
            loader_state["mode"] = f"in-process (worker pool failed: {type(e).__name__})"
            plan_futs.append(None)

    def plans_for(fut, vb, qb):
        if fut is not None:
            try:
                return PlanPool.attach(fut, vb, qb)
            except Exception as e:                          # noqa: BLE001
                loader_state["pool"] = None
                loader_state["mode"] = f"in-process (worker pool failed: {type(e).__name__})"
        return attach_plan(vb), attach_plan(qb, kind="txt")

    stage_t = collections.defaultdict(float)

    def stage(i, total):
        t0 = time.perf_counter()
        vb, qb = host[i % n_host]
        vb, qb = plans_for(plan_futs.popleft(), dict(vb), dict(qb))
        t1 = time.perf_counter()
        if i + 2 < total:
            submit_plan(i + 2)
        t2 = time.perf_counter()
        stage_t["plan_wait"] += t1 - t0
        stage_t["plan_submit"] += t2 - t1
        (vb_dev, qb_dev), ev, slot = stager.stage(vb, qb)    # preallocated ring, side stream
        stage_t["h2d_enqueue"] += time.perf_counter() - t2
        return vb_dev, qb_dev, ev, slot

    result_host = torch.zeros(2, dtype=torch.float32).pin_memory()

    def e2e_loop(n):
        """Every step: wait for its prefetched inputs, enqueue fwd+bwd, start the (async) D2H
        read of a result scalar, then build the NEXT batch's plan on the host and start its H2D
        copy on the side stream while the GPU computes. The scalar of step i is consumed right
        after step i+1 has been enqueued (one-step-lagged logging), so the host never idles the
        GPU; all n results are read."""
        out, pending = 0.0, None
        dbg = True     # host phase times go to stderr (diagnostics; the JSON line stays on stdout)
        tt = collections.defaultdict(float)
        dones = []
        nxt = stage(0, n)
        for i in range(n):
            t_a = time.perf_counter()
            vb_dev, qb_dev, ev, ring_slot = nxt
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            record_plans((vb_dev, qb_dev), cur)   # allocator safety across streams
            gflat.zero_()
            t_b = time.perf_counter()
            clip = fwd_bwd(vb_dev, qb_dev)
            t_c = time.perf_counter()
            slot = result_host[i % 2:i % 2 + 1]
            slot.copy_(clip[0, 0, :8].float().sum().reshape(1), non_blocking=True)   # D2H
            done = torch.cuda.Event(enable_timing=True)
            done.record(cur)
            dones.append(done)
            stager.release(ring_slot, cur)
            t_d = time.perf_counter()
            if i + 1 < n:
                nxt = stage(i + 1, n)                   # plan upload + H2D overlap this step's compute
            t_e = time.perf_counter()
            if pending is not None:
                pending[0].synchronize()
                out += float(pending[1][0])
            pending = (done, slot)
            t_f = time.perf_counter()
            for k, v in (("record", t_b - t_a), ("fwd_bwd", t_c - t_b), ("d2h", t_d - t_c),
                         ("stage", t_e - t_d), ("sync_prev", t_f - t_e)):
                tt[k] += v
        pending[0].synchronize()
        out += float(pending[1][0])
        if dbg:
            print("e2e stage() totals ms:", {k: round(v * 1e3, 1) for k, v in stage_t.items()},
                  file=sys.stderr)
            stage_t.clear()
            gaps = [round(a.elapsed_time(b), 2) for a, b in zip(dones[:-1], dones[1:])]
            print("e2e device ms between step ends:", gaps, file=sys.stderr)
            print("e2e host phases ms/step:", {k: round(v / n * 1e3, 2) for k, v in tt.items()},
                  file=sys.stderr)
        return out

    def prime():     # a running loader always has two batches of plans in flight
        plan_futs.clear()
        submit_plan(0)
        submit_plan(1)

    # start every loader worker (each imports torch once) before anything is timed
    if loader_state["pool"] is not None:
        try:
            for f in [loader_state["pool"].submit(*host[0]) for _ in range(6)]:
                f.result(timeout=300)
        except Exception as e:                              # noqa: BLE001
            loader_state["pool"] = None
            loader_state["mode"] = f"in-process (worker pool failed: {type(e).__name__})"
    prime()
    e2e_loop(max(3, args.warmup))
    prime()
    for f in plan_futs:
        if f is not None:
            try:
                f.result(timeout=300)
            except Exception:                               # noqa: BLE001
                pass
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if loader_state["pool"] is not None:
        loader_state["pool"].shutdown()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    bi = h2d_bytes(host[0][0]) + h2d_bytes(host[0][1])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(sample_clips=args.cpu_clips, steps=2, warmup=1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "SYN-TVR-dense: HierarchicalVlModel 'repr' + CrossModalTrm 'txt' "
                                   "fwd+bwd (hero_finetune dims 6+3 layers), per-rank 32 clips x "
                                   "100 frames x 4352-d + 640 rows x (5 frames + 20 tokens) + 32 "
                                   "queries x 16 tokens",
                       "global_batch": world * B, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "dropout": 0.1, "optimizer_in_step": bool(args.with_optimizer),
                       "query_rows": "separate call" if args.separate_txt else
                       "fused into the video-row pass (forward_repr_txt)",
                       "allreduce_in_step": world > 1 and not args.dp_skip_exchange,
                       "allreduce_overlap": (("per-layer buckets during backward, "
                                              + ("peer copies over NVLink (copy engines)"
                                                 if bucketer.p2p is not None else "NCCL")
                                              + "; remainder NCCL after backward")
                                             if bucketer is not None else
                                             "none: one NCCL all-reduce of the flat gradient "
                                             "buffer after backward" if world > 1 else "none"),
                       "e2e_plans": loader_state["mode"],
                       "l2": "no explicit flush: per-step working set (~0.35 GB weights+grads, "
                             "~3 GB activations, 111 MB inputs) exceeds the 126 MB L2"},
            "clocks": clocks, "gpu_launches": launches,
            "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
            "e2e": {"value": round(e2e_value, 2), "unit": "clips/s", "h2d_bytes_per_step": bi,
                    "d2h_bytes_per_step": 4},
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_baseline(sample_clips=4, steps=2, warmup=1):
    """The oracle port of the reference path on the host cores: fwd+bwd (autograd) on a bounded
    sample of the same workload (first `sample_clips` clips of SYN-TVR-dense)."""
    from hero_b200 import synth
    from oracle import hero_oracle as orc
    torch.manual_seed(0)
    P = orc.seeded_weights(orc.param_shapes(), seed=0)
    P = {k: v.requires_grad_(True) for k, v in P.items()}
    vb, qb = synth.syn_tvr_dense(batch_size=sample_clips, seed=1234)
    g = torch.Generator().manual_seed(7)
    dclip = torch.randn(sample_clips, 100, H, generator=g) * 1e-2
    dq = torch.randn(sample_clips, qb["input_ids"].shape[1], H, generator=g) * 1e-2

    def step():
        for v in P.values():
            v.grad = None
        clip = orc.hierarchical_repr(P, vb, F_LAYERS, C_LAYERS, HEADS)
        q = orc.cross_modal_txt(P, "f_encoder.", qb, F_LAYERS, HEADS)
        torch.autograd.backward([clip, q], [dclip, dq])

    for _ in range(warmup):
        step()
    best = float("inf")
    t_all = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        t_all += dt
    return {"value": round(sample_clips / (t_all / steps), 3), "unit": "clips/s",
            "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
            "best_value": round(sample_clips / best, 3),
            "sample": f"oracle fwd+bwd (fp32, torch CPU autograd) on the first {sample_clips} "
                      f"clips of SYN-TVR-dense, {steps} timed steps after {warmup} warm-up"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU path, restated by the oracle port (the Python
    reference cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    cb = cpu_baseline(sample_clips=args.cpu_clips, steps=steps, warmup=warm)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "clips/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": round(1e3 * args.cpu_clips / cb["value"], 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SYN-TVR-dense sample of {args.cpu_clips} clips per step "
                                   "(same per-clip shapes as the GPU arm)", "parallelism": "cpu"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--with-optimizer", action="store_true")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N>1: force one all-reduce of the flat gradient after backward (already "
                         "the default; overrides --dp-transport)")
    ap.add_argument("--dp-transport", default="none", choices=("none", "p2p", "nccl", "auto"),
                    help="N>1: 'none' = one NCCL all-reduce of the flat gradient buffer after "
                         "backward (default, measured fastest end to end); 'p2p' / 'nccl' = "
                         "GradBucketer: buckets travel during backward")
    ap.add_argument("--dp-skip-exchange", action="store_true",
                    help="DIAGNOSTIC (invalid as a result): N>1 without any gradient exchange, to "
                         "separate per-GPU compute time from communication")
    ap.add_argument("--bucket-elems", type=int, default=1 << 20,
                    help="N>1: gradient ranges are exchanged once this many elements are final")
    ap.add_argument("--overlap-ctas", type=int, default=0,
                    help="N>1: CTAs of the communicator that exchanges gradient buckets during "
                         "backward (the compute kernels leave that many SMs free)")
    ap.add_argument("--separate-txt", action="store_true",
                    help="encode the query rows with a separate f_encoder(batch, 'txt') call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=4)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
