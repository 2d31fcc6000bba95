# round 2, session 3: scatter CTA shape A/B (512 threads x 8192-key tile x 2 CTAs/SM  vs  256 x 4096 x 4  vs  1024 x 8192 x 2)
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
for v in base sc256 sc1024; do
if [ $v = base ]; then unset MC_B200_LIB; else export MC_B200_LIB=$PWD/metacherchant_b200/csrc/variants/$v/libmcgpu.so; fi
$B --steps 3 --warmup 2 > gpurun_out/r2z_$v.log 2> gpurun_out/r2z_$v.err
python -c "
import json,sys
f='gpurun_out/r2z_$v.log'
d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'])" || tail -3 gpurun_out/r2z_$v.err
done
