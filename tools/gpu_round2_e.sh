set -x
timeout 300 python -m pytest tests/test_gpu_mlp.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_gputest_mlp.log 2>&1
tail -15 gpurun_out/r2_gputest_mlp.log
timeout 300 python tools/mlp_bench.py > gpurun_out/r2_mlp_bench.log 2>&1
cat gpurun_out/r2_mlp_bench.log
timeout 600 python -m pytest tests/test_gpu_al_loop.py tests/test_gpu_datagen.py tests/test_gpu_shim.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest4.log 2>&1
tail -5 gpurun_out/r2_gputest4.log
timeout 300 python tools/shim_latency.py > gpurun_out/r2_shim_latency.log 2>&1
cat gpurun_out/r2_shim_latency.log
timeout 600 python tools/pipeline_device_bench.py 3 1024 4096 > gpurun_out/r2_pipeline_device2.log 2>&1
cat gpurun_out/r2_pipeline_device2.log
