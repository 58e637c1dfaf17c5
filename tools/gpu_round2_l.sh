set -x
timeout 600 python -m pytest tests/test_gpu_drivers.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest_drv.log 2>&1
tail -4 gpurun_out/r2_gputest_drv.log
timeout 900 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
tail -c 1500 gpurun_out/r2_bench_d.json; tail -3 gpurun_out/r2_bench_d.err
timeout 1500 python tools/certify.py --c4 16384 --c23 8192 --c5 8192 > gpurun_out/r2_certify_large.md 2> gpurun_out/r2_certify_large.err
head -12 gpurun_out/r2_certify_large.md | cut -c1-250
timeout 900 python tools/agreement.py 8192 32768 > gpurun_out/r2_agreement_large.md 2> gpurun_out/r2_agreement_large.err
head -20 gpurun_out/r2_agreement_large.md | cut -c1-250
