#!/bin/bash
# round 2, GPU call 17: staging chunk size A/B on the C2 e2e, ramped slices on the C3 e2e
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline"
for c in 1048576 2097152 4194304; do
  S2G_STAGE_CHUNK=$c timeout 600 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2r_c2_chunk$c.json 2> gpurun_out/r2r_c2_chunk$c.err
done
timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2r_c3.json 2> gpurun_out/r2r_c3.err
timeout 300 python -m pytest tests/test_gpu_parity_3d_healpix.py -q -m gpu -x -k "3d" > gpurun_out/r2r_tests.log 2>&1; tail -n 2 gpurun_out/r2r_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "%.1f ms"%d["ms_per_step"], "e2e", round(e.get("ms_per_step"),1), "pinned", round(e.get("pinned_ms_per_step"),1), e.get("phases_last_call"))
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
