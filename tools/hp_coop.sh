# timing of the HEALPix cooperative split (S2G_HP_COOP_RINGS) on the C4 samples; run under gpurun
for r in ${RINGS:-128 192 256}; do
  S2G_HP_COOP_RINGS=$r python bench.py --workload c4s --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4s coop=$r', round(d['ms_per_step'],1),'ms')"
done
for r in ${FULL:-256}; do
  S2G_HP_COOP_RINGS=$r python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 coop=$r', round(d['ms_per_step'],1),'ms', round(d['value'],2), d['unit'])"
done
