#!/bin/bash
# round 2, GPU call 30: pass A of the 2D gather class by one thread per particle where the closed form applies — C2, the
# C5 sample, 2D parity tests (with the new ordered-scatter / ordered-stencil tests)
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
timeout 600 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2C_c2.json 2> gpurun_out/r2C_c2.err
timeout 600 $B --workload c5s --steps 3 --warmup 2 > gpurun_out/r2C_c5s.json 2> gpurun_out/r2C_c5s.err
timeout 1200 python -m pytest tests -q -m gpu -x -k "2d or 2D or golden or baseline or fp32 or sedov or stokes or stencil or closed or block" > gpurun_out/r2C_tests.log 2>&1; tail -n 3 gpurun_out/r2C_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2C_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
