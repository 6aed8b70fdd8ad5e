#!/bin/bash
# round 2, GPU call 1: the new parity tests (extended-precision HEALPix arbiter, BASELINE streams), then the whole suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_baseline_streams.py tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py -q -m gpu -x --durations=15 > gpurun_out/r2_new_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_new_tests.log
timeout 1500 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/r2_all_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_all_tests.log
tail -5 gpurun_out/r2_new_tests.log gpurun_out/r2_all_tests.log
# the rewritten bench.py: quick sanity on the small workload, then the reference arm sanity (CPU)
timeout 600 python bench.py --workload small --steps 2 --warmup 1 --extra none > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err
echo "bench small rc=$?"; tail -c 600 gpurun_out/r2_bench_small.err
timeout 900 python bench.py --workload c4s --steps 1 --warmup 1 --extra none > gpurun_out/r2_bench_c4s.json 2> gpurun_out/r2_bench_c4s.err
echo "bench c4s rc=$?"; tail -c 600 gpurun_out/r2_bench_c4s.err
