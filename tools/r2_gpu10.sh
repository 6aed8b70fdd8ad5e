#!/bin/bash
# round 2, GPU call 10: 3D AUTO threshold sweep on C3, ncu --set full of the two HEALPix gather passes (small-angle list)
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline"
for c in 2048 4096 8192 16384; do
  S2G_GATHER3D_MIN_CELLS=$c timeout 600 $B --workload c3 --steps 2 --warmup 2 --no-e2e > gpurun_out/r2j_c3_min$c.json 2> gpurun_out/r2j_c3_min$c.err
done
S2G_3D_CACHE=0 timeout 600 $B --workload c3 --steps 2 --warmup 2 --no-e2e > gpurun_out/r2j_c3_nocache.json 2> gpurun_out/r2j_c3_nocache.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hp_gather -s 4 -c 2 -o gpurun_out/r2_prof_hpgatherAB -f $B --workload c4s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2j_ncu_hpg.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "pairs", d["config"]["pairs"])
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
ls -la gpurun_out/r2_prof_hpgatherAB.ncu-rep
