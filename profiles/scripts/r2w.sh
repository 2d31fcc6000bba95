# round 2, super-k-mer front end: full GPU suite, bench with the front end on and off, trace of one step
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2w_pytest.log; tail -3 gpurun_out/r2w_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
MC_TRACE=1 $B --steps 1 --warmup 1 > gpurun_out/r2w_trace.log 2> gpurun_out/r2w_trace.err; grep -c "sk expand" gpurun_out/r2w_trace.err; grep "sk expand" gpurun_out/r2w_trace.err | head -12
$B --steps 3 --warmup 2 > gpurun_out/r2w_bench_sk.log 2> gpurun_out/r2w_bench_sk.err
python -c "
import json
for f in ('gpurun_out/r2w_bench_sk.log',):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], r.get('host_ms_per_step'))"
tail -3 gpurun_out/r2w_bench_sk.err
