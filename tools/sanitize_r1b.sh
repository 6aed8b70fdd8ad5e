#!/bin/bash
# compute-sanitizer over the kernels added late in round 1 (ordered Stokes compositing, cooperative HEALPix launch,
# fused projections); run under gpurun.
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_stokes_rm.py -x -q -m gpu -k "planes_and_dtypes or order_dependence or quirks" 2>&1 | tail -3
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_3d_healpix.py -x -q -k "cooperative" 2>&1 | tail -3
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_projection.py -x -q -m gpu -k "float32" 2>&1 | tail -3
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_stokes_rm.py -x -q -m gpu -k "order_dependence or quirks" 2>&1 | tail -3
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_3d_healpix.py -x -q -k "cooperative" 2>&1 | tail -3
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_3d_healpix.py -x -q -k "cooperative" 2>&1 | tail -3
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_stokes_rm.py -x -q -m gpu -k "quirks or order_dependence" 2>&1 | tail -3
