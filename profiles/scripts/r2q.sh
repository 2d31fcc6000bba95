# round 2, capture "q": BFS with the visited-set read overlapped with the table probe (204 environment tests), bin planning by density
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env.py tests/test_gpu_cli.py tests/test_gpu_dist.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2q_pytest.log; tail -3 gpurun_out/r2q_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 4 --warmup 2"
MC_TRACE=1 $B > gpurun_out/r2q_bench.log 2> gpurun_out/r2q_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2q_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('r2q', d['ms_per_step'], r['frac'], r['stage_ms_per_step']); print('  env', d['env']['env_ms'], d['env']['reads_to_env_s'])"
grep -E "k_bfs|environments -> files|bfs passes" gpurun_out/r2q_bench.err | tail -4
$B --force-sharded --no-env > gpurun_out/r2q_bench_sharded1.log 2> gpurun_out/r2q_bench_sharded1.err; python -c "
import json
d=json.loads(open('gpurun_out/r2q_bench_sharded1.log').read().strip().splitlines()[-1]); r=d['roofline']; print('sharded world=1', d['ms_per_step'], d['phases_ms_last_step'], r['stage_ms_per_step'])"
tail -2 gpurun_out/r2q_bench_sharded1.err
