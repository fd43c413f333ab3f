#!/bin/bash
# last validation visit of the round: full GPU suite, parity printout, smoke(), default bench
mkdir -p gpurun_out
TAG=${1:-run}
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_bench_path_gpu.py -m gpu -q -s --no-header -p no:cacheprovider 2>&1 | grep -E "clip outputs|query rows|worst gradient|pre-temporal|passed|failed" > gpurun_out/${TAG}_parity.txt
cat gpurun_out/${TAG}_parity.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
python - <<PY
import json
txt=open("gpurun_out/${TAG}_bench.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "host", d["host_enqueue_ms_per_step"], d.get("host_enqueue_ms_queue_not_full"), "frac", d["roofline"]["frac"], d["roofline"]["step_frac_of_peak"], "launches", d["gpu_launches"], d["clocks"])
print("gpu_ref", (d.get("gpu_reference") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
