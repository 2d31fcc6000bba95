mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
for i in 1 2; do
$B --steps 4 --warmup 2 > gpurun_out/r3g_$i.log 2> gpurun_out/r3g_$i.err
python -c "
import json,sys
f='gpurun_out/r3g_$i.log'
d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], d['verify'])" || tail -3 gpurun_out/r3g_$i.err
done
