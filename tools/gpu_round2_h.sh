set -x
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest_full.log 2>&1
tail -6 gpurun_out/r2_gputest_full.log
timeout 900 python tools/certify.py > gpurun_out/r2_certify.md 2> gpurun_out/r2_certify.err
tail -3 gpurun_out/r2_certify.err
timeout 600 python bench.py --steps 2 --warmup 3 --cpu-pipeline 0 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-pipeline 0 > gpurun_out/r2_ncu_bench.log 2>&1
tail -c 600 gpurun_out/r2_bench_c.json
grep -c solve_kernel gpurun_out/r2_bench_launches.csv
