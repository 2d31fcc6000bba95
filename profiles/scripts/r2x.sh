mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_superkmers.py -x -q 2>&1 | tail -5
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
# (the MC_SK_NOEMIT experiment hook -- emission compiled out -- was removed from the library after this measurement)
MC_SK_NOEMIT=1 $B --steps 2 --warmup 1 > gpurun_out/r2x_noemit.log 2> gpurun_out/r2x_noemit.err
python -c "
import json
for f in ('gpurun_out/r2x_noemit.log',):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], r.get('host_ms_per_step'))"
tail -3 gpurun_out/r2x_noemit.err
