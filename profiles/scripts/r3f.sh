# session 3: L2 bulk prefetch of the next aggregation unit (nopf = without), A/B on one box
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
for v in pf nopf pf nopf; do
if [ $v = pf ]; then unset MC_B200_LIB; else export MC_B200_LIB=$PWD/metacherchant_b200/csrc/variants/$v/libmcgpu.so; fi
$B --steps 4 --warmup 2 > gpurun_out/r3f_$v.log 2> gpurun_out/r3f_$v.err
python -c "
import json,sys
f='gpurun_out/r3f_$v.log'
d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'])" || tail -3 gpurun_out/r3f_$v.err
done
