#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-run}
timeout 1200 python -m pytest tests/test_heads_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_heads.log 2>&1
echo "heads rc=$?" >> gpurun_out/${TAG}_heads.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
tail -30 gpurun_out/${TAG}_heads.log | cut -c1-300; python - <<PY
import json
txt=open("gpurun_out/${TAG}_bench.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"], json.dumps(d["extra"])[:1500])
PY
tail -3 gpurun_out/${TAG}_bench.err | cut -c1-300
