# AGG_FU = 1 adopted: parity of the counting path, smoke, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py tests/test_gpu_superkmers.py -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/r3w_bench.log 2> gpurun_out/r3w_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r3w_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], 'allocs', d['driver_allocs_in_timed_region'], 'e2e', d['e2e']['value'], 'env', d['env']['env_ms'], d['env']['reads_to_env_s'], d['verify'])"
