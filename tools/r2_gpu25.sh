#!/bin/bash
# round 2, GPU call 25: scatter lists of the 2D deposit in 64x64-pixel block order — tiny workload with the order on / off,
# C2 and the C5 sample unchanged?, 2D parity tests
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
timeout 600 $B --workload tiny --steps 3 --warmup 2 > gpurun_out/r2x_tiny_order.json 2> gpurun_out/r2x_tiny_order.err
S2G_2D_ORDER=0 timeout 600 $B --workload tiny --steps 3 --warmup 2 > gpurun_out/r2x_tiny_noorder.json 2> gpurun_out/r2x_tiny_noorder.err
timeout 600 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2x_c2.json 2> gpurun_out/r2x_c2.err
timeout 600 $B --workload c5s --steps 3 --warmup 2 > gpurun_out/r2x_c5s.json 2> gpurun_out/r2x_c5s.err
S2G_2D_ORDER=0 timeout 600 $B --workload c5s --steps 3 --warmup 2 > gpurun_out/r2x_c5s_noorder.json 2> gpurun_out/r2x_c5s_noorder.err
timeout 900 python -m pytest tests -q -m gpu -x -k "2d or 2D or golden or tiny or baseline or fp32 or sedov or stokes" > gpurun_out/r2x_tests.log 2>&1; tail -n 2 gpurun_out/r2x_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2x_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
