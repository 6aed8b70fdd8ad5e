#!/bin/bash
# round 2, GPU call 18: staging through own pinned bounce buffers — the staging test, then e2e of C2 / C3 / C4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_2d.py -q -m gpu -x -k "staging" > gpurun_out/r2s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_tests.log; tail -n 3 gpurun_out/r2s_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2s_c2.json 2> gpurun_out/r2s_c2.err
S2G_STAGE_BOUNCE=0 timeout 600 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2s_c2_nobounce.json 2> gpurun_out/r2s_c2_nobounce.err
timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2s_c3.json 2> gpurun_out/r2s_c3.err
timeout 900 $B --workload c4 --steps 1 --warmup 1 > gpurun_out/r2s_c4.json 2> gpurun_out/r2s_c4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "%.1f ms"%d["ms_per_step"], "e2e", round(e.get("ms_per_step"),1), "pinned", e.get("pinned_ms_per_step"), e.get("phases_last_call"))
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
