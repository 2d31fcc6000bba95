mkdir -p gpurun_out
python bench.py --force-sharded --no-e2e --no-cpu-baseline --no-env --steps 3 --warmup 2 > gpurun_out/r3o_sharded1.log 2> gpurun_out/r3o_sharded1.err
python -c "
import json
d=json.loads(open('gpurun_out/r3o_sharded1.log').read().strip().splitlines()[-1]); print('force-sharded N=1', d['ms_per_step'], d['phases_ms_last_step'], d['verify'])" || tail -5 gpurun_out/r3o_sharded1.err
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
