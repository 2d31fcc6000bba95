# session 3: aggregation flush depth (bucket probes in flight per thread: 1 / 2 / 3) and shared-table probe limit (6 / 12)
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
for v in base fu3 fu1 pr6; do
if [ $v = base ]; then unset MC_B200_LIB; else export MC_B200_LIB=$PWD/metacherchant_b200/csrc/variants/$v/libmcgpu.so; fi
$B --steps 3 --warmup 2 > gpurun_out/r3v_$v.log 2> gpurun_out/r3v_$v.err
python -c "
import json,sys
f='gpurun_out/r3v_$v.log'
d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, d['ms_per_step'], r['stage_ms_per_step'])" || tail -3 gpurun_out/r3v_$v.err
done
