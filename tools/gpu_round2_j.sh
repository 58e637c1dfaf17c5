set -x
timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_al_loop.py tests/test_gpu_mpc.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest_guess.log 2>&1
tail -15 gpurun_out/r2_gputest_guess.log
timeout 300 python tools/prof_run.py 37888 3 | tail -1
timeout 300 python tools/config_bench.py 2>&1 | tail -8
