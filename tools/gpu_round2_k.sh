set -x
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_gputest_final.log 2>&1
tail -4 gpurun_out/r2_gputest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
tail -c 900 gpurun_out/r2_bench_ref_final.json
timeout 1200 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 1800 gpurun_out/r2_bench_final.json
