# round 2, capture "h": clock sampler A/B on one box (NVML thread vs nvidia-smi child vs none)
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 6 --warmup 3"
run() { name=$1; shift; $B "$@" > gpurun_out/r2h_$name.log 2> gpurun_out/r2h_$name.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h_$name.log').read().strip().splitlines()[-1]); print('$name', d['ms_per_step'], 'stage sum', d['roofline']['stage_sum_ms_per_step'], d['clocks'])"; tail -2 gpurun_out/r2h_$name.err; }
run nvml
run smi --clocks smi
run none --no-clocks
run nvml2
run smi2 --clocks smi
run none2 --no-clocks
