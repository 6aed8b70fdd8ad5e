#!/bin/bash
# usage (under gpurun --gpus 2): tools/r1_gpu2.sh — device group over real peers, FP32-mode tests and diagnostics
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus2.txt 2>&1
timeout 300 python -m pytest tests/test_device_group.py tests/test_fp32_accumulate.py -q -m gpu --durations=5 \
    > gpurun_out/gpu2_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu2_tests.log
tail -4 gpurun_out/gpu2_tests.log
timeout 100 python tools/fp32_diag.py > gpurun_out/fp32_diag.jsonl 2> gpurun_out/fp32_diag.err
cat gpurun_out/fp32_diag.jsonl | cut -c1-400
S2G_GROUP_BENCH_N=16777216 S2G_GROUP_BENCH_NPIX=4096 timeout 200 python tools/group_bench.py \
    > gpurun_out/group_bench_c2.json 2> gpurun_out/group_bench_c2.err
tail -c 900 gpurun_out/group_bench_c2.json
S2G_GROUP_NO_P2P=1 S2G_GROUP_BENCH_N=16777216 S2G_GROUP_BENCH_NPIX=4096 timeout 200 python tools/group_bench.py \
    > gpurun_out/group_bench_c2_nop2p.json 2> gpurun_out/group_bench_c2_nop2p.err
tail -c 400 gpurun_out/group_bench_c2_nop2p.json
