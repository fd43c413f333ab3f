#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-run}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
echo "rc=$?" >> gpurun_out/${TAG}_bench2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-dp-overlap --no-gpu-reference --no-extra > gpurun_out/${TAG}_bench2_noov.json 2> gpurun_out/${TAG}_bench2_noov.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --dp-wire fp32 --no-dp-overlap --no-gpu-reference --no-extra > gpurun_out/${TAG}_bench2_fp32.json 2> gpurun_out/${TAG}_bench2_fp32.err
for f in bench2 bench2_noov bench2_fp32; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$f.json"))
    print("$f", d["value"], d["ms_per_step"], d["median_ms_per_step"], d["e2e"]["value"], d.get("gpu_reference"), d["config"]["allreduce"], d["config"]["allreduce_check"])
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -5 gpurun_out/${TAG}_bench2.err
