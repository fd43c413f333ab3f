#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-run}
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "attention or column_sums" > gpurun_out/${TAG}_attn.log 2>&1
echo "attn rc=$?" >> gpurun_out/${TAG}_attn.log
tail -25 gpurun_out/${TAG}_attn.log | cut -c1-400
timeout 300 python tools/attn_bench.py > gpurun_out/${TAG}_attn_bench.txt 2>&1; cat gpurun_out/${TAG}_attn_bench.txt | tail -5
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
python - <<PY
import json
txt=open("gpurun_out/${TAG}_bench.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], d.get("host_enqueue_ms_queue_not_full"), d["roofline"]["frac"], d["gpu_launches"])
print(json.dumps(d["extra"])[:1500])
PY
