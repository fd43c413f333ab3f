#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-run}
timeout 900 ncu --set full --import-source on --clock-control none -o gpurun_out/${TAG}_gemm python tools/profile_gemm.py > gpurun_out/${TAG}_ncu.log 2>&1
timeout 600 python tools/host_profile.py > gpurun_out/${TAG}_host_profile.txt 2>&1
STEP_TIMELINE=gpurun_out/${TAG}_timeline.csv timeout 600 python tools/step_profile.py 5 > gpurun_out/${TAG}_step_profile.md 2>&1
ls -la gpurun_out | tail -8; tail -50 gpurun_out/${TAG}_step_profile.md
