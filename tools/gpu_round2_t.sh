set -x
# final evidence of round 2 on the final library (2-warp CTAs in the batch kernel)
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2t_gputest.log 2>&1
tail -4 gpurun_out/r2t_gputest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2t_smoke.log 2>&1; tail -1 gpurun_out/r2t_smoke.log
timeout 1200 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -c 300 gpurun_out/r2t_bench.json; tail -3 gpurun_out/r2t_bench.err
timeout 600 python tools/prof_run.py 37888 3 > gpurun_out/r2t_prof_plain.log 2>&1 && tail -2 gpurun_out/r2t_prof_plain.log
timeout 600 python tools/prof_run.py 1 20 > gpurun_out/r2t_prof_lone.log 2>&1 && tail -2 gpurun_out/r2t_prof_lone.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-pipeline 0 --extras 0 --pipeline 0 --cpu-sample 16 > gpurun_out/r2t_ncu_bench.log 2>&1
tail -c 200 gpurun_out/r2t_ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -o gpurun_out/prof_r2t python tools/prof_run.py 37888 3 > gpurun_out/r2t_ncu_a.log 2>&1
tail -2 gpurun_out/r2t_ncu_a.log
