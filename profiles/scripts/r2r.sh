# round 2, capture "r": BFS lookahead prefetch by the idle lanes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
MC_TRACE=1 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r2r_bench.log 2> gpurun_out/r2r_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2r_bench.log').read().strip().splitlines()[-1]); print('  env', d['env']['env_ms'], d['env']['reads_to_env_s'])"
grep -E "k_bfs|environments -> files|bfs passes" gpurun_out/r2r_bench.err | tail -4
