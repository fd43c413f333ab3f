#!/bin/bash
# N-GPU bench (trimmed extras) ; usage: gpu_dp8.sh TAG N [extra bench flags]
mkdir -p gpurun_out
TAG=${1:-run}; N=${2:-8}; shift; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-pretrain-mix --roofline-steps 3 "$@" > gpurun_out/${TAG}_bench${N}.json 2> gpurun_out/${TAG}_bench${N}.err
echo "rc=$?" >> gpurun_out/${TAG}_bench${N}.err
python - <<PY
import json
try:
    txt=open("gpurun_out/${TAG}_bench${N}.json").read()
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("N=$N", d["value"], d["ms_per_step"], d.get("median_ms_per_step"), "e2e", d["e2e"]["value"], "gpu_ref", d.get("gpu_reference"), d["config"]["allreduce"], d["config"]["allreduce_check"], "host", d["host_enqueue_ms_per_step"], d.get("host_enqueue_ms_queue_not_full"))
    print(json.dumps(d.get("extra", {}))[:800])
except Exception as e:
    print("FAILED", e)
PY
tail -5 gpurun_out/${TAG}_bench${N}.err
