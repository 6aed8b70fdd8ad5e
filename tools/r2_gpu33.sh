#!/bin/bash
# round 2, GPU call 33: result maps copied back by the staging threads (unstage_output) — its test, the staged-path test,
# the whole GPU suite with the thresholds lowered so that most calls take the threaded copies both ways, e2e of C3 / C2,
# ncu --set full of k_scatter3d on the block-ordered class list (c3s)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "staging or result_map" > gpurun_out/r2E_tests_new.log 2>&1; tail -n 2 gpurun_out/r2E_tests_new.log
S2G_UNSTAGE_MIN=4096 S2G_STAGE_MIN=2000 timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2E_tests_all_lowered.log 2>&1; tail -n 3 gpurun_out/r2E_tests_all_lowered.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 300 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2E_c3.json 2> gpurun_out/r2E_c3.err
timeout 300 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2E_c2.json 2> gpurun_out/r2E_c2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_scatter3d -s 1 -c 1 -f -o gpurun_out/r2E_k_scatter3d python bench.py --workload c3s --steps 1 --warmup 1 --extra none --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2E_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2E_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), "e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned_ms_per_step")})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
