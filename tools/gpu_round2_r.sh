# CTA shape A/B on the BENCH (retirement granularity matters when launches overlap): 4 / 2 / 1 warps per CTA, 20 warps per SM
for v in base w2 w1 base w1; do
  echo "== $v"
  VBOC_LIB=$PWD/vboc_b200/variants/$v.so timeout 600 python tools/prof_run.py 37888 3 2>&1 | tail -1
  VBOC_LIB=$PWD/vboc_b200/variants/$v.so timeout 900 python bench.py --steps 4 --warmup 4 --extras 0 --pipeline 0 --cpu-sample 16 > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2r_bench_$v.json').read().strip().splitlines()[-1])
    print('$v', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1))
except Exception as e:
    print('$v failed', e); print(open('gpurun_out/r2r_bench_$v.err').read()[-600:])
PY
done
