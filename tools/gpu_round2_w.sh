set -x
( time timeout 860 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench_k20.json 2> gpurun_out/r2w_bench_k20.err ) 2> gpurun_out/r2w_time.log
tail -3 gpurun_out/r2w_time.log
tail -c 400 gpurun_out/r2w_bench_k20.json; tail -2 gpurun_out/r2w_bench_k20.err
