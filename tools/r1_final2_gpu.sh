#!/bin/bash
# usage (under gpurun, one GPU): end-of-round check — the whole GPU suite, smoke(), the default bench line (C2, with
# e2e and the CPU baseline)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_gpu_tests.log; tail -3 gpurun_out/final_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -3 gpurun_out/final_smoke.log
timeout 200 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
tail -c 1500 gpurun_out/final_bench_c2.json
