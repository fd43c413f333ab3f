#!/bin/bash
# Final round-2 profile artifacts: ncu launch list of bench steps, ncu --set full digests of the
# attention / LayerNorm kernels and of the GEMM epilogue families (digested ON the box: reports with
# source are too large to bring back), CUPTI step timeline.
mkdir -p gpurun_out
TAG=${1:-r02f}
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 -c 560 --csv --log-file gpurun_out/${TAG}_launches_bench_step.csv python bench.py --steps 3 --warmup 5 --no-extra --no-pretrain-mix --no-gpu-reference --no-cpu-baseline --roofline-steps 2 > gpurun_out/${TAG}_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_bench_step.csv --step > gpurun_out/${TAG}_launches_bench_step.md 2>&1
tail -3 gpurun_out/${TAG}_launches_bench_step.md
cat > /tmp/prof_misc.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from hero_b200 import ops
dev = torch.device("cuda:0")
M, H, heads = 16512, 768, 12
x32 = torch.randn(M, H, device=dev); dy = torch.randn(M, H, device=dev).bfloat16()
g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
y = torch.empty(M, H, dtype=torch.bfloat16, device=dev); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
dx = torch.empty_like(y); dxd = torch.empty_like(y); dg, db, dbi = (torch.zeros(H, device=dev) for _ in range(3))
from hero_b200.plan import SeqPlan, DeviceIndex
import numpy as np
sp = SeqPlan(np.ones((M // 25, 25), np.int64)); di = DeviceIndex(sp.arrays("s_"), dev); att = sp.attn(di, "s_")
n = sp.n_tok
qkv = torch.randn(n, 3 * H, device=dev).bfloat16(); ctx = torch.empty(n, H, dtype=torch.bfloat16, device=dev)
lse = torch.empty(n, heads, device=dev); dctx = torch.randn(n, H, device=dev).bfloat16(); dqkv = torch.empty_like(qkv)
dbq = torch.zeros(3 * H, device=dev)
drop = ops.drop_params(0.1, 5)
for _ in range(2):
    ops.ln_fwd(x32, g, b, 1e-12, y, n_rows=M, mean=mean, rstd=rstd)
    ops.ln_bwd(dy, x32, g, mean, rstd, n_rows=M, dx=dx, dx_drop=dxd, drop2=drop, dgamma=dg, dbeta=db, dbias=dbi)
    ops.attn_fwd(qkv, att, ctx, heads=heads, drop=drop, lse=lse)
    ops.attn_bwd(qkv, att, ctx, dctx, lse, dqkv, heads=heads, drop=drop, dbias=dbq)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --import-source on --clock-control none -o /tmp/${TAG}_misc python /tmp/prof_misc.py > gpurun_out/${TAG}_ncu_misc.log 2>&1
python tools/ncu_top.py /tmp/${TAG}_misc.ncu-rep gpurun_out/${TAG}_misc_full_top.txt 14 > gpurun_out/${TAG}_ncu_top.log 2>&1
REPS=1 timeout 1200 ncu --set full --import-source on --clock-control none -o /tmp/${TAG}_gemm python tools/profile_gemm.py > gpurun_out/${TAG}_ncu_gemm.log 2>&1
python tools/ncu_top.py /tmp/${TAG}_gemm.ncu-rep gpurun_out/${TAG}_gemm_full_top.txt 18 >> gpurun_out/${TAG}_ncu_top.log 2>&1
timeout 600 python tools/step_profile.py 5 > gpurun_out/${TAG}_step_profile.txt 2>&1
ls -la gpurun_out | grep ${TAG}_ | tail -12; grep -n "==== kernel\|time_duration\|stalls:" gpurun_out/${TAG}_misc_full_top.txt | head -40
