#!/bin/bash
# round 2, GPU call 13: asin variant only above 0.2 rad; 3D deposit overlapped with the pageable staging
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py tests/test_gpu_baseline_streams.py tests/test_device_group.py -q -m gpu -x > gpurun_out/r2m_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2m_tests.log; tail -n 3 gpurun_out/r2m_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2m_c4s.json 2> gpurun_out/r2m_c4s.err
timeout 1200 $B --workload c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2m_c4.json 2> gpurun_out/r2m_c4.err
timeout 900 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2m_c3.json 2> gpurun_out/r2m_c3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "e2e", e.get("ms_per_step"), e.get("pinned_ms_per_step"), e.get("phases_last_call"))
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
