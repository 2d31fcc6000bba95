# round 2, capture "i": host synchronisation policy A/B on one box (spin vs the driver's default), three runs each, interleaved
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 6 --warmup 3"
run() { name=$1; shift; "$@" > gpurun_out/r2i_$name.log 2> gpurun_out/r2i_$name.err; python -c "
import json
d=json.loads(open('gpurun_out/r2i_$name.log').read().strip().splitlines()[-1]); print('$name', d['ms_per_step'], 'stage sum', d['roofline']['stage_sum_ms_per_step'], d['clocks'].get('how'))"; tail -2 gpurun_out/r2i_$name.err; }
nproc; uptime
for i in 1 2 3; do
run spin$i $B
run auto$i env MC_SYNC=auto $B
done
