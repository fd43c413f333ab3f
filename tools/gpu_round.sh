#!/bin/bash
# One GPU-box visit: kernel tests, encoder parity tests, bench. Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-run}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/${TAG}_kernels.log 2>&1
echo "kernels rc=$?" >> gpurun_out/${TAG}_kernels.log
timeout 2400 python -m pytest tests/test_encoder_gpu.py tests/test_bench_path_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/${TAG}_encoder.log 2>&1
echo "encoder rc=$?" >> gpurun_out/${TAG}_encoder.log
HERO_GEMM_PROFILE_DUMP=gpurun_out/${TAG}_gemm_per_launch.csv timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_kernels.log; tail -30 gpurun_out/${TAG}_encoder.log; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
