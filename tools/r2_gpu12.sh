#!/bin/bash
# round 2, GPU call 12: parked band ranges in the pair expansion; launch list of the full C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py -q -m gpu -x -k "healpix or golden" > gpurun_out/r2l_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2l_tests.log; tail -n 3 gpurun_out/r2l_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2l_c4s.json 2> gpurun_out/r2l_c4s.err
timeout 1200 $B --workload c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2l_c4.json 2> gpurun_out/r2l_c4.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2l_launches_c4.csv $B --workload c4 --steps 1 --warmup 0 --no-e2e > gpurun_out/r2l_launches_c4.log 2>&1
python - <<'PY'
import json,glob,csv,collections
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "pairs", d["config"]["pairs"])
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
try:
    rows=[r for r in csv.reader(open("gpurun_out/r2l_launches_c4.csv")) if len(r)>5]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
    agg=collections.OrderedDict()
    for r in rows[1:]:
        v=float(r[vi].replace(",","")); v = v/1e6 if r[ui]=="ns" else (v/1e3 if r[ui]=="us" else v)
        k=r[ki][:64]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
    tot=sum(t for n,t in agg.values())
    for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]: print("  %-64s n=%4d %10.2f ms %5.1f%%"%(k,n,t,100*t/tot))
except Exception as ex: print("launch list ERR", ex)
PY
