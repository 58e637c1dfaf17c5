set -x
timeout 900 python -m pytest tests/test_gpu_mpc.py -x -q --tb=short -p no:cacheprovider > gpurun_out/r2o_gputest_mpc.log 2>&1
tail -15 gpurun_out/r2o_gputest_mpc.log
