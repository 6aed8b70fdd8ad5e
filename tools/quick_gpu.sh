#!/bin/bash
# usage (under gpurun): tools/quick_gpu.sh <tag>  -> 2D parity tests, C2 bench phases, ncu of k_gather2d on "small"
tag=$1
python -m pytest tests/test_gpu_parity_2d.py -x -q 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_c2_$tag.log
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c2_$tag.log").read())
print("C2:", round(d["value"],3), "Mp/s", round(d["ms_per_step"],1), "ms", {k: round(v,1) for k,v in d["roofline"]["fp64"]["phase_ms"].items()}, d["config"]["pairs"])
PY
if [ "$2" != "noncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:${3:-k_gather2d} -c 1 -o gpurun_out/prof_$tag -f python bench.py --workload small --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
fi
