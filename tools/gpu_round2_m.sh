set -x
timeout 900 python -m pytest tests/test_gpu_datagen.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest_tg.log 2>&1
tail -6 gpurun_out/r2_gputest_tg.log
timeout 900 python bench.py --steps 2 --warmup 3 --cpu-pipeline 0 --pipeline 256 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
tail -c 1300 gpurun_out/r2_bench_e.json; tail -3 gpurun_out/r2_bench_e.err
