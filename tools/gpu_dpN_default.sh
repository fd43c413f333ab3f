#!/bin/bash
# the driver's invocation at N GPUs (default flags), both arms ; usage: gpu_dpN_default.sh TAG N
mkdir -p gpurun_out
TAG=${1:-run}; N=${2:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench${N}.json 2> gpurun_out/${TAG}_bench${N}.err
echo "rc=$?" >> gpurun_out/${TAG}_bench${N}.err
python - <<PY
import json
try:
    txt=open("gpurun_out/${TAG}_bench${N}.json").read()
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("N=$N", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "gpu_ref", (d.get("gpu_reference") or {}).get("value"), d["config"]["allreduce_check"], "cpu", d.get("cpu_baseline"))
    print(json.dumps(d.get("extra", {}))[:1800])
except Exception as e:
    print("FAILED", e)
PY
tail -3 gpurun_out/${TAG}_bench${N}.err
