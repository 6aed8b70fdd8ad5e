#!/bin/bash
# usage (under gpurun, one GPU): the ncu launch list of the bench command (shares of the step per kernel), FP64 mode
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_v2.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_c2_v2.log 2>&1
wc -l gpurun_out/launches_c2_v2.csv
