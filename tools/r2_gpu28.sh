#!/bin/bash
# round 2, GPU call 28: scatter kernels with 32 particle records per warp visit (shared-memory record board) and the
# CIC/TSC stencils in block order — tiny, c3cic / c3tsc with the order on / off, C2, 2D + stencil parity tests
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
timeout 600 $B --workload tiny --steps 3 --warmup 2 > gpurun_out/r2A_tiny.json 2> gpurun_out/r2A_tiny.err
timeout 600 $B --workload c3cic --steps 3 --warmup 2 > gpurun_out/r2A_c3cic.json 2> gpurun_out/r2A_c3cic.err
S2G_STENCIL_ORDER=0 timeout 600 $B --workload c3cic --steps 3 --warmup 2 > gpurun_out/r2A_c3cic_noorder.json 2> gpurun_out/r2A_c3cic_noorder.err
timeout 600 $B --workload c3tsc --steps 3 --warmup 2 > gpurun_out/r2A_c3tsc.json 2> gpurun_out/r2A_c3tsc.err
S2G_STENCIL_ORDER=0 timeout 600 $B --workload c3tsc --steps 3 --warmup 2 > gpurun_out/r2A_c3tsc_noorder.json 2> gpurun_out/r2A_c3tsc_noorder.err
timeout 600 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2A_c2.json 2> gpurun_out/r2A_c2.err
timeout 900 python -m pytest tests -q -m gpu -x -k "2d or 2D or golden or tiny or baseline or fp32 or sedov or stokes or stencil or cic or tsc or gadget" > gpurun_out/r2A_tests.log 2>&1; tail -n 2 gpurun_out/r2A_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2A_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
