#!/bin/bash
# usage (under gpurun, one GPU): tools/r1_final_gpu.sh
# new features first (device group, FP32-accumulate mode), then the whole GPU suite, then the FP32-mode bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_device_group.py tests/test_fp32_accumulate.py tests -q -m gpu --durations=12 \
    > gpurun_out/gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
timeout 150 python bench.py --steps 2 --warmup 3 --accum f32 --no-cpu-baseline --no-e2e \
    > gpurun_out/bench_c2_f32.json 2> gpurun_out/bench_c2_f32.err
tail -c 600 gpurun_out/bench_c2_f32.json
timeout 120 python tools/group_bench.py > gpurun_out/group_bench.json 2> gpurun_out/group_bench.err
tail -c 600 gpurun_out/group_bench.json
