set -x
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/r2y_gputest.log 2>&1
tail -4 gpurun_out/r2y_gputest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
