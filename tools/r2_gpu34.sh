#!/bin/bash
# round 2, GPU call 34 (last): the final build — smoke, the whole GPU suite at default thresholds, and the host-array
# tests with the thresholds lowered so that positions and maps go back through the staging threads
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2F_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2F_smoke.log
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2F_all_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2F_all_tests.log; tail -n 3 gpurun_out/r2F_all_tests.log
S2G_UNSTAGE_MIN=4096 S2G_STAGE_MIN=2000 timeout 200 python -m pytest tests -q -m gpu -k "end_to_end or staging or result_map or healpix_map or sort_z or projection or device_group or group" > gpurun_out/r2F_tests_lowered.log 2>&1; tail -n 2 gpurun_out/r2F_tests_lowered.log
