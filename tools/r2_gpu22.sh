#!/bin/bash
# round 2, GPU call 22: final validation (tools/r2_gpu16.sh) + a full ncu capture of every k_hp_gather launch of one c4s step
bash tools/r2_gpu16.sh
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hp_gather -s 6 -c 6 -f -o gpurun_out/r2p_k_hp_gather python bench.py --workload c4s --steps 1 --warmup 1 --extra none --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2p_ncu_hp.log 2>&1
ls -la gpurun_out/r2p_k_hp_gather.ncu-rep
