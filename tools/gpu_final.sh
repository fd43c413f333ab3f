#!/bin/bash
# final validation visit: full GPU test suite, smoke(), microbenchmarks, the driver's default bench
mkdir -p gpurun_out
TAG=${1:-run}
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -6 gpurun_out/${TAG}_tests.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 300 python tools/ln_bench.py > gpurun_out/${TAG}_ln_bench.txt 2>&1; tail -6 gpurun_out/${TAG}_ln_bench.txt
timeout 300 python tools/attn_bench.py > gpurun_out/${TAG}_attn_bench.txt 2>&1; tail -3 gpurun_out/${TAG}_attn_bench.txt
timeout 600 python tools/gemm_bench.py > gpurun_out/${TAG}_gemm_bench.txt 2>&1; tail -30 gpurun_out/${TAG}_gemm_bench.txt
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python - <<PY
import json
txt=open("gpurun_out/${TAG}_bench.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "host", d["host_enqueue_ms_per_step"], d.get("host_enqueue_ms_queue_not_full"), "frac", d["roofline"]["frac"], d["roofline"]["step_frac_of_peak"], "launches", d["gpu_launches"], d["clocks"])
print("gpu_ref", d.get("gpu_reference"), "cpu", d.get("cpu_baseline"))
print(json.dumps(d["extra"])[:1200])
print(open("gpurun_out/${TAG}_bench_ref.json").read()[-700:])
PY
