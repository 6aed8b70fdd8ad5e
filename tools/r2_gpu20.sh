#!/bin/bash
# round 2, GPU call 20: discs over a pole through the tile-gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py -q -m gpu -x > gpurun_out/r2t_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2t_tests.log; tail -n 3 gpurun_out/r2t_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2t_c4s.json 2> gpurun_out/r2t_c4s.err
timeout 1200 $B --workload c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2t_c4.json 2> gpurun_out/r2t_c4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "pairs", d["config"]["pairs"])
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
