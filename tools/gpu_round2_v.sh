set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --pipeline 512 > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.err
tail -c 1500 gpurun_out/r2v_bench_n2.json; tail -3 gpurun_out/r2v_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2v_bench_ref_n2.json 2> gpurun_out/r2v_bench_ref_n2.err
tail -c 300 gpurun_out/r2v_bench_ref_n2.json
