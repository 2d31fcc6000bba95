# the final tree on cfg4 (200 M reads, k = 23, 500 targets)
mkdir -p gpurun_out
python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r3y_cfg4.log 2> gpurun_out/r3y_cfg4.err
python -c "
import json
d=json.loads(open('gpurun_out/r3y_cfg4.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg4', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], d['driver_allocs_in_timed_region'], d['env']['env_ms'], d['env']['targets_at_cap'])" || tail -5 gpurun_out/r3y_cfg4.err
