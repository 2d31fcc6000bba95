# round 2, capture "o": lockstep merge kernel on one GPU (world = 1 runs the whole N > 1 path: local count, owner dump, merge, filter)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 4 --warmup 2"
$B --force-sharded > gpurun_out/r2o_bench_sharded1.log 2> gpurun_out/r2o_bench_sharded1.err; python -c "
import json
d=json.loads(open('gpurun_out/r2o_bench_sharded1.log').read().strip().splitlines()[-1]); r=d['roofline']; print('sharded world=1', d['ms_per_step'], d['phases_ms_last_step'], r['stage_ms_per_step'], d['verify'])"
tail -2 gpurun_out/r2o_bench_sharded1.err
