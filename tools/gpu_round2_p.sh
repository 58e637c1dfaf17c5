set -x
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r2p_gputest.log 2>&1
tail -8 gpurun_out/r2p_gputest.log
