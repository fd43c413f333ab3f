#!/bin/bash
# A/B of two bench flag sets on ONE box, interleaved: gpu_ab.sh TAG "<flags A>" "<flags B>"
mkdir -p gpurun_out
TAG=${1:-ab}; A="$2"; B="$3"
: > gpurun_out/${TAG}.txt
for rep in 1 2 3; do
  for v in A B; do
    if [ $v = A ]; then F="$A"; else F="$B"; fi
    line=$(timeout 300 python bench.py --steps 30 --warmup 5 --no-extra --no-gpu-reference --no-cpu-baseline --no-pretrain-mix --roofline-steps 3 $F 2>/dev/null | grep '^{' | tail -1)
    python - "$v" "$F" "$line" >> gpurun_out/${TAG}.txt <<'PY'
import json, sys
v, f, line = sys.argv[1:4]
try:
    d = json.loads(line)
    print(f"{v} [{f}] ms_per_step {d['ms_per_step']:.4f} median {d.get('median_ms_per_step')} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} clocks {d['clocks'].get('sm_mhz')}")
except Exception as ex:
    print(v, "failed", ex)
PY
  done
done
cat gpurun_out/${TAG}.txt
