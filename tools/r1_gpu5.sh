#!/bin/bash
# usage (under gpurun, one GPU): A/B of k_gather2d CTA geometries (prebuilt alternative libraries
# sphtogrid.jl_b200/libs2g_alt_*.so, see S2G_G2D_* in csrc/s2g_gather2d.cu) on the C2 step, same box, then the 2D
# parity tests on the fastest one
mkdir -p gpurun_out
cd sphtogrid.jl_b200 && cp libsphtogrid_cuda.so libs2g_alt_default.so && cd ..
best=default; best_ms=1000000
for v in default a b c; do
  cp sphtogrid.jl_b200/libs2g_alt_$v.so sphtogrid.jl_b200/libsphtogrid_cuda.so
  timeout 100 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/geom_$v.json 2> gpurun_out/geom_$v.err
  ms=$(python -c "
import json; d=json.load(open('gpurun_out/geom_$v.json')); print(round(d['ms_per_step'],1), {k: round(x,1) for k,x in d['roofline']['fp64']['phase_ms'].items()}, d['config']['pairs'])")
  echo "variant $v: $ms"
  m=$(python -c "import json; print(int(json.load(open('gpurun_out/geom_$v.json'))['ms_per_step']*10))")
  if [ "$m" -lt "$best_ms" ]; then best_ms=$m; best=$v; fi
done
echo "best: $best"
cp sphtogrid.jl_b200/libs2g_alt_$best.so sphtogrid.jl_b200/libsphtogrid_cuda.so
timeout 120 python -m pytest tests/test_gpu_parity_2d.py tests/test_fp32_accumulate.py tests/test_golden_vectors.py -q -m gpu > gpurun_out/geom_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/geom_tests.log; tail -4 gpurun_out/geom_tests.log
timeout 60 python bench.py --steps 2 --warmup 3 --accum f32 --no-cpu-baseline --no-e2e > gpurun_out/geom_${best}_f32.json 2> /dev/null
python -c "
import json; d=json.load(open('gpurun_out/geom_${best}_f32.json')); print('f32 mode on $best:', round(d['ms_per_step'],1))"
