#!/bin/bash
# round 2, GPU call 26: ncu --set full of the two scatter kernels on the tiny sample (4 Mi particles of the tiny stream)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scatter2d -s 2 -c 2 -f -o gpurun_out/r2y_k_scatter2d_tiny python bench.py --workload tinys --steps 1 --warmup 1 --extra none --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2y_ncu.log 2>&1
ls -la gpurun_out/r2y_k_scatter2d_tiny.ncu-rep; tail -n 3 gpurun_out/r2y_ncu.log
