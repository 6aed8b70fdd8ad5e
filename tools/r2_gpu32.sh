#!/bin/bash
# round 2, GPU call 32: evidence for the final build — launch list of a C4 step, ncu --set full of the two 2D scatter
# kernels on the tiny sample
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2D_launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 1 --extra none --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2D_launches_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scatter2d -s 2 -c 2 -f -o gpurun_out/r2D_k_scatter2d_tiny python bench.py --workload tinys --steps 1 --warmup 1 --extra none --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2D_ncu.log 2>&1
ls -la gpurun_out/r2D_*
