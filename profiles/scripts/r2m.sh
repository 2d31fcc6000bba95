# round 2, capture "m": parser with one mask plane + flags from bytes in pass 2; run-per-thread merge (world = 1 profile of the N > 1 path)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 8 --warmup 3"
$B > gpurun_out/r2m_bench.log 2> gpurun_out/r2m_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('r2m', d['ms_per_step'], r['frac'], 'stage sum', r['stage_sum_ms_per_step'], r['stage_ms_per_step'])"
tail -2 gpurun_out/r2m_bench.err
$B --force-sharded --steps 4 > gpurun_out/r2m_bench_sharded1.log 2> gpurun_out/r2m_bench_sharded1.err; python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench_sharded1.log').read().strip().splitlines()[-1]); r=d['roofline']; print('sharded world=1', d['ms_per_step'], d['phases_ms_last_step'], r['stage_ms_per_step'], d['verify'])"
tail -2 gpurun_out/r2m_bench_sharded1.err
