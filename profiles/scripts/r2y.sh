# round 2, session 3: k_bin_agg lane loop A/B + parity of the counting path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
for v in 1 0 1; do
$B --steps 4 --warmup 2 --tune agg_lane_loop=$v > gpurun_out/r2y_ll$v.log 2> gpurun_out/r2y_ll$v.err
python -c "
import json,sys
f='gpurun_out/r2y_ll$v.log'
d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'])"
done
