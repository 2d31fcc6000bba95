# round 2, capture "d": 2 x B200 -- multi-GPU parity (table, exchange digests, gathered solid set, target-sharded environments) and the N = 2 bench
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests/test_gpu_dist.py tests/test_gpu_env.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2d_pytest_n$N.log; tail -4 gpurun_out/r2d_pytest_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/dist_parity_worker.py > gpurun_out/r2d_dist_parity_n$N.log 2>&1; grep -E "DIST_|Error|error" gpurun_out/r2d_dist_parity_n$N.log | head
MC_DIST_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2d_bench_n$N.log 2> gpurun_out/r2d_bench_n$N.err; grep '"metric"' gpurun_out/r2d_bench_n$N.log | tail -1 | cut -c1-6000; tail -5 gpurun_out/r2d_bench_n$N.err
