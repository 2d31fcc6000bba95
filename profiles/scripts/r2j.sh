# round 2, capture "j": reads-classifier parity + host-side phase times of the counting step (three runs)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_classify.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2j_pytest.log; tail -8 gpurun_out/r2j_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 8 --warmup 3"
for i in 1 2 3; do
$B > gpurun_out/r2j_bench$i.log 2> gpurun_out/r2j_bench$i.err; python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench$i.log').read().strip().splitlines()[-1]); r=d['roofline']; print('run$i', d['ms_per_step'], 'stage sum', r['stage_sum_ms_per_step'], 'walls', r['step_wall_ms']); print('   host', r['host_ms_per_step']); print('   dev ', r['stage_ms_per_step'])"
tail -2 gpurun_out/r2j_bench$i.err
done
