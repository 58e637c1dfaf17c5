set -x
timeout 600 python -m pytest tests/test_gpu_mpc.py tests/test_gpu_parity.py tests/test_gpu_shim.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest_mpc.log 2>&1
tail -25 gpurun_out/r2_gputest_mpc.log
timeout 300 python tools/prof_run.py 37888 3 | tail -1
