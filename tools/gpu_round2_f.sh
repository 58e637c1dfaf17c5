set -x
timeout 300 python -m pytest tests/test_gpu_mlp.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_gputest_mlp2.log 2>&1
tail -4 gpurun_out/r2_gputest_mlp2.log
timeout 300 python tools/mlp_bench.py > gpurun_out/r2_mlp_bench2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_pipe_kernel -s 1 -c 1 -o gpurun_out/prof_r2_mlp python tools/mlp_bench.py > gpurun_out/r2_ncu_mlp.log 2>&1
cat gpurun_out/r2_mlp_bench2.log
tail -2 gpurun_out/r2_ncu_mlp.log
