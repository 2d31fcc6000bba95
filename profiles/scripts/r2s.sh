# round 2, capture "s": extractor staging by cp.async.bulk + mbarrier; owner dump staged in shared memory (world = 1 profile)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2s_pytest.log; tail -3 gpurun_out/r2s_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-env --steps 6 --warmup 3"
for t in 1 0; do
$B --tune extract_tma=$t > gpurun_out/r2s_bench_tma$t.log 2> gpurun_out/r2s_bench_tma$t.err; python -c "
import json
d=json.loads(open('gpurun_out/r2s_bench_tma$t.log').read().strip().splitlines()[-1]); r=d['roofline']; print('extract_tma=$t', d['ms_per_step'], r['frac'], r['stage_ms_per_step'])"
tail -2 gpurun_out/r2s_bench_tma$t.err
done
$B --force-sharded --steps 3 > gpurun_out/r2s_bench_sharded1.log 2> gpurun_out/r2s_bench_sharded1.err; python -c "
import json
d=json.loads(open('gpurun_out/r2s_bench_sharded1.log').read().strip().splitlines()[-1]); r=d['roofline']; print('sharded world=1', d['ms_per_step'], d['phases_ms_last_step'], d['verify'] and d['verify']['ok'])"
tail -2 gpurun_out/r2s_bench_sharded1.err
