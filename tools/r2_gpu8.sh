#!/bin/bash
# round 2, GPU call 8: compile-time series length (NT 5 / 8 lists), slice-size A/B on the full C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py tests/test_gpu_baseline_streams.py -q -m gpu -x -k "healpix or golden" > gpurun_out/r2h_hp_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2h_hp_tests.log; tail -n 3 gpurun_out/r2h_hp_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2h_c4s.json 2> gpurun_out/r2h_c4s.err
timeout 1200 $B --workload c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2h_c4.json 2> gpurun_out/r2h_c4.err
S2G_HP_BATCH_PARTICLES=16777216 S2G_PAIR_CAP=1200000000 timeout 1200 $B --workload c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2h_c4_b16m.json 2> gpurun_out/r2h_c4_b16m.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches_c4s.csv $B --workload c4s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2h_launches_c4s.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_hp_gather.3, 0, 0, 5" -c 1 -o gpurun_out/r2_prof_hpgatherB -f $B --workload c4s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2h_ncu_hpgB.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_hp_gather.3, 0, 1, 5" -c 1 -o gpurun_out/r2_prof_hpgatherA -f $B --workload c4s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2h_ncu_hpgA.log 2>&1
python - <<'PY'
import json,glob,csv,collections
for f in sorted(glob.glob("gpurun_out/r2h_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "pairs", d["config"]["pairs"])
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
try:
    rows=[r for r in csv.reader(open("gpurun_out/r2h_launches_c4s.csv")) if len(r)>5]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
    agg=collections.OrderedDict()
    for r in rows[1:]:
        v=float(r[vi].replace(",","")); v = v/1e6 if r[ui]=="ns" else (v/1e3 if r[ui]=="us" else v)
        k=r[ki][:64]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
    for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print("  %-64s n=%3d %9.2f ms"%(k,n,t))
except Exception as ex: print("launch list ERR", ex)
PY
ls -la gpurun_out/r2_prof_hpgather[AB].ncu-rep
