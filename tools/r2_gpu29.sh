#!/bin/bash
# round 2, GPU call 29: host->device staging on 4 helper threads (S2G_STAGE_THREADS) — e2e of C2 / C3 with 1 and 4
# threads, e2e of C4 with 4, the staged-path tests
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 python -m pytest tests -q -m gpu -x -k "staging or staged or e2e or host" > gpurun_out/r2B_tests.log 2>&1; tail -n 2 gpurun_out/r2B_tests.log
for w in c2 c3; do
  timeout 600 $B --workload $w --steps 3 --warmup 2 > gpurun_out/r2B_${w}_t4.json 2> gpurun_out/r2B_${w}_t4.err
  S2G_STAGE_THREADS=1 timeout 600 $B --workload $w --steps 3 --warmup 2 > gpurun_out/r2B_${w}_t1.json 2> gpurun_out/r2B_${w}_t1.err
  S2G_STAGE_THREADS=8 timeout 600 $B --workload $w --steps 3 --warmup 2 > gpurun_out/r2B_${w}_t8.json 2> gpurun_out/r2B_${w}_t8.err
done
timeout 900 $B --workload c4 --steps 1 --warmup 1 > gpurun_out/r2B_c4_t4.json 2> gpurun_out/r2B_c4_t4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2B_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), "e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned_ms_per_step","phases")})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
