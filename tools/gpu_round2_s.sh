set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py tests/test_gpu_stream.py tests/test_gpu_cartesian.py tests/test_gpu_datagen.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2s_gputest.log 2>&1
tail -4 gpurun_out/r2s_gputest.log
timeout 600 python tools/prof_run.py 37888 3 2>&1 | tail -1
timeout 900 python bench.py --steps 4 --warmup 4 --extras 0 --pipeline 0 --cpu-sample 16 > gpurun_out/r2s_bench4.json 2> gpurun_out/r2s_bench4.err
python -c "
import json
d=json.loads(open('gpurun_out/r2s_bench4.json').read().strip().splitlines()[-1])
print('K=4 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1))"
