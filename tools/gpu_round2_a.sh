set -x
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest1.log 2>&1
tail -5 gpurun_out/r2_gputest1.log
python tools/certify.py > gpurun_out/r2_certify.md 2> gpurun_out/r2_certify.err
python tools/agreement.py 2048 8192 > gpurun_out/r2_agreement.md 2> gpurun_out/r2_agreement.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2> gpurun_out/r2_bench_ref_a.err
python tools/prof_run.py 8192 8 > gpurun_out/r2_prof_base.log 2>&1
tail -3 gpurun_out/r2_prof_base.log
