#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-run}
timeout 1500 python -m pytest tests/test_heads_gpu.py tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_heads.log 2>&1
echo "heads rc=$?" >> gpurun_out/${TAG}_heads.log
timeout 2000 python -m pytest tests/test_encoder_gpu.py tests/test_bench_path_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_encoder.log 2>&1
echo "encoder rc=$?" >> gpurun_out/${TAG}_encoder.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
grep -E "passed|failed|Error|error" gpurun_out/${TAG}_heads.log | tail -15 | cut -c1-300; tail -4 gpurun_out/${TAG}_encoder.log | cut -c1-300; python - <<PY
import json
txt=open("gpurun_out/${TAG}_bench.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["host_enqueue_ms_per_step"], json.dumps(d["extra"].get("pretrain_mix"))[:600])
PY
