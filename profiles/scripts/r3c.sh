mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
for i in 1 2; do
$B --steps 4 --warmup 3 > gpurun_out/r3c_bench$i.log 2> gpurun_out/r3c_bench$i.err
python -c "
import json
d=json.loads(open('gpurun_out/r3c_bench$i.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], d['driver_allocs_in_timed_region'])"
done
for w in cfg5 cfg4; do
python bench.py --workload $w --no-e2e --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r3c_$w.log 2> gpurun_out/r3c_$w.err
python -c "
import json
d=json.loads(open('gpurun_out/r3c_$w.log').read().strip().splitlines()[-1]); r=d['roofline']; print('$w', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], d['driver_allocs_in_timed_region'], d.get('env'))" || tail -5 gpurun_out/r3c_$w.err
done
