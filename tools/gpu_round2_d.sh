set -x
timeout 600 python -m pytest tests/test_gpu_datagen.py tests/test_gpu_certify.py tests/test_gpu_parity.py tests/test_gpu_drivers.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest3.log 2>&1
tail -8 gpurun_out/r2_gputest3.log
timeout 600 python tools/pipeline_device_bench.py 3 1024 8192 32768 > gpurun_out/r2_pipeline_device.log 2>&1
timeout 600 python tools/pipeline_device_bench.py 2 1024 8192 >> gpurun_out/r2_pipeline_device.log 2>&1
cat gpurun_out/r2_pipeline_device.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -c 1500 gpurun_out/r2_bench_b.json
timeout 600 python tools/prof_run.py 37888 3 > gpurun_out/r2_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -o gpurun_out/prof_r2a python tools/prof_run.py 37888 3 > gpurun_out/r2_ncu_a.log 2>&1
tail -3 gpurun_out/r2_ncu_a.log
