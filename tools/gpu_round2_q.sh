set -x
# final evidence of round 2: full GPU suite, smoke, both bench arms, launch list of the bench command, bounded ncu capture
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2q_gputest.log 2>&1
tail -4 gpurun_out/r2q_gputest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2q_smoke.log 2>&1; tail -1 gpurun_out/r2q_smoke.log
timeout 1200 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
tail -c 600 gpurun_out/r2q_bench.json; tail -3 gpurun_out/r2q_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err
tail -c 300 gpurun_out/r2q_bench_ref.json
timeout 600 python tools/prof_run.py 37888 3 > gpurun_out/r2q_prof_plain.log 2>&1 && tail -2 gpurun_out/r2q_prof_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-pipeline 0 --extras 0 --pipeline 0 > gpurun_out/r2q_ncu_bench.log 2>&1
tail -2 gpurun_out/r2q_ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -o gpurun_out/prof_r2q python tools/prof_run.py 37888 3 > gpurun_out/r2q_ncu_a.log 2>&1
tail -2 gpurun_out/r2q_ncu_a.log
