#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run under gpurun).  Round 1: 0 errors / 0 hazards.
set -e
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_2d.py -x -q -k "large_footprints or multi_image or empty or closed_form or non_finite or sphmapping_end" 2>&1 | tail -4
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_3d_healpix.py -x -q -k "deposit_3d_parity or all_regimes or filter_sort or stencils" 2>&1 | tail -4
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_2d.py -x -q -k "large_footprints or closed_form" 2>&1 | tail -3
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_3d_healpix.py -x -q -k "all_regimes or resolved" 2>&1 | tail -3
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_3d_healpix.py -x -q -k "all_regimes" 2>&1 | tail -3
