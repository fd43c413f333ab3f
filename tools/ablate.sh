#!/bin/bash
# What each kernel family costs in the real (two-stream, PDL-overlapped) step: bench.py with the
# family skipped inside the layer stack. Needs tools/build_ablate.py run first (CPU box).
# bits: 0 colsum, 1 attn fwd, 2 attn bwd, 3 LN fwd, 4 LN bwd, 5 wgrad GEMMs, 6 dgrad GEMMs, 7 fwd GEMMs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export HERO_B200_LIB=$PWD/hero_b200/libhero_b200_ablate.so
OUT=gpurun_out/${1:-ablate}.txt
: > $OUT
for mask in 0 1 2 4 6 8 16 24 32 64 128 224 0; do
  line=$(HERO_ABLATE=$mask timeout 300 python bench.py --steps 30 --warmup 5 --no-extra --no-gpu-reference \
         --no-cpu-baseline --no-pretrain-mix --roofline-steps 3 2>/dev/null | grep '^{' | tail -1)
  python - "$mask" "$line" >> $OUT <<'PY'
import json, sys
mask, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"mask {int(mask):4d}  ms_per_step {d['ms_per_step']:.3f}  value {d['value']:.1f}  "
          f"e2e {d['e2e']['value']:.1f}  clocks {d['clocks'].get('sm_mhz')}")
except Exception as ex:
    print(f"mask {mask}: failed ({ex}) {line[:200]}")
PY
done
cat $OUT
