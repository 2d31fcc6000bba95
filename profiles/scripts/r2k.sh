# round 2, capture "k": step time spread after taking cudaMemGetInfo out of the step (four runs)
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 8 --warmup 3"
for i in 1 2 3 4; do
$B > gpurun_out/r2k_bench$i.log 2> gpurun_out/r2k_bench$i.err; python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench$i.log').read().strip().splitlines()[-1]); r=d['roofline']; print('run$i', d['ms_per_step'], 'stage sum', r['stage_sum_ms_per_step'], 'walls', r['step_wall_ms']); print('   host', r['host_ms_per_step'])"
tail -2 gpurun_out/r2k_bench$i.err
done
