set -x
# sanity of the committed state: full GPU suite, smoke, both bench arms
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r2n_gputest.log 2>&1
tail -5 gpurun_out/r2n_gputest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2n_smoke.log 2>&1; tail -2 gpurun_out/r2n_smoke.log
timeout 1200 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -c 2500 gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err
tail -c 800 gpurun_out/r2n_bench_ref.json; tail -3 gpurun_out/r2n_bench_ref.err
