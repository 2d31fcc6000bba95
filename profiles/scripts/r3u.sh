# last validation of the committed tree: full GPU suite + the N > 1 code path on one rank
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3u_pytest.log; tail -3 gpurun_out/r3u_pytest.log
python bench.py --force-sharded --no-e2e --no-cpu-baseline --no-env --steps 3 --warmup 2 > gpurun_out/r3u_sharded1.log 2> gpurun_out/r3u_sharded1.err
python -c "
import json
d=json.loads(open('gpurun_out/r3u_sharded1.log').read().strip().splitlines()[-1]); print('force-sharded N=1', d['ms_per_step'], d['phases_ms_last_step'], d['verify'] and d['verify']['ok'])" || tail -5 gpurun_out/r3u_sharded1.err
