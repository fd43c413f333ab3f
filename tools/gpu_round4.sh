#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-run}
timeout 1500 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/${TAG}_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/${TAG}_kernels.log
timeout 2400 python -m pytest tests/test_encoder_gpu.py tests/test_bench_path_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/${TAG}_encoder.log 2>&1
echo "encoder rc=$?" >> gpurun_out/${TAG}_encoder.log
timeout 600 python tools/ln_bench.py > gpurun_out/${TAG}_ln_bench.txt 2>&1
timeout 600 python tools/host_profile.py > gpurun_out/${TAG}_host_profile.txt 2>&1
STEP_TIMELINE=gpurun_out/${TAG}_timeline.csv timeout 600 python tools/step_profile.py 5 > gpurun_out/${TAG}_step_profile.md 2>&1
HERO_GEMM_PROFILE_DUMP=gpurun_out/${TAG}_gemm_per_launch.csv timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_kernels.log; grep -E "clip outputs:|query rows:|worst gradient|pre-temporal|passed|failed" gpurun_out/${TAG}_encoder.log | cut -c1-300; cat gpurun_out/${TAG}_ln_bench.txt; head -30 gpurun_out/${TAG}_step_profile.md; tail -3 gpurun_out/${TAG}_step_profile.md; cat gpurun_out/${TAG}_bench.json | cut -c1-600
