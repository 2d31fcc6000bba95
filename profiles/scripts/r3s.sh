# session 3, last validation of the committed tree: full GPU suite, smoke, default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3s_pytest.log; tail -3 gpurun_out/r3s_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
python bench.py > gpurun_out/r3s_bench.log 2> gpurun_out/r3s_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r3s_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], 'allocs', d['driver_allocs_in_timed_region'], 'e2e', d['e2e']['value'], 'env', d['env']['env_ms'], d['env']['reads_to_env_s'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('env_s'), d['clocks'])"
tail -2 gpurun_out/r3s_bench.err
python bench.py --force-sharded --no-e2e --no-cpu-baseline --no-env --steps 3 --warmup 2 > gpurun_out/r3s_sharded1.log 2> gpurun_out/r3s_sharded1.err
python -c "
import json
d=json.loads(open('gpurun_out/r3s_sharded1.log').read().strip().splitlines()[-1]); print('force-sharded N=1', d['ms_per_step'], d['phases_ms_last_step'], d['verify'] and d['verify']['ok'])" || tail -5 gpurun_out/r3s_sharded1.err
