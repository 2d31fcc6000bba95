# round 2, capture "g": where does the host gap (ms_per_step - sum of stages) come from?  same box, back to back
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 6 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/r2g_$name.log 2> gpurun_out/r2g_$name.err; python -c "
import json
d=json.loads(open('gpurun_out/r2g_$name.log').read().strip().splitlines()[-1]); print('$name', d['ms_per_step'], 'stage sum', d['roofline']['stage_sum_ms_per_step'], 'launches', d['gpu_launches'], d['clocks'])"; }
run default
run noclocks --no-clocks
run notimers --tune stage_timing=0
run noclocks_notimers --no-clocks --tune stage_timing=0
run default2
