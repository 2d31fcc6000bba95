mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r3d_cfg4.log 2> gpurun_out/r3d_cfg4.err
python -c "
import json
d=json.loads(open('gpurun_out/r3d_cfg4.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg4', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], d['driver_allocs_in_timed_region'], d['env']['env_ms'])" || tail -5 gpurun_out/r3d_cfg4.err
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
$B --steps 4 --warmup 3 > gpurun_out/r3d_bench.log 2> gpurun_out/r3d_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3d_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], d['driver_allocs_in_timed_region'])"
