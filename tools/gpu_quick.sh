#!/bin/bash
# quick visit: selected kernel tests (-k "$2"), attention microbenchmark, short bench
mkdir -p gpurun_out
TAG=${1:-run}
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "${2:-attention}" > gpurun_out/${TAG}_k.log 2>&1
echo "kernels rc=$?" >> gpurun_out/${TAG}_k.log
tail -8 gpurun_out/${TAG}_k.log | cut -c1-300
timeout 300 python tools/attn_bench.py > gpurun_out/${TAG}_attn_bench.txt 2>&1; tail -4 gpurun_out/${TAG}_attn_bench.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-pretrain-mix --roofline-steps 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
python - <<PY
import json
txt=open("gpurun_out/${TAG}_bench.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], d.get("host_enqueue_ms_queue_not_full"), d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
PY
