# round 2, super-k-mer front end: first parity run + a quick bench with stage times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_superkmers.py -x -q 2>&1 | tail -25 > gpurun_out/r2v_sk.log; tail -25 gpurun_out/r2v_sk.log
