set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/al_sharded_check.py > gpurun_out/r2_al_sharded.log 2>&1
tail -4 gpurun_out/r2_al_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 2500 gpurun_out/r2_bench_n2.json
tail -3 gpurun_out/r2_bench_n2.err
