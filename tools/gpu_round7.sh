#!/bin/bash
# full GPU test suite + default bench + family ablations
mkdir -p gpurun_out
TAG=${1:-run}
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
python - <<PY
import json
txt=open("gpurun_out/${TAG}_bench.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], d.get("host_enqueue_ms_queue_not_full"), d["roofline"]["frac"])
PY
bash tools/ablate.sh ${TAG}_ablate
