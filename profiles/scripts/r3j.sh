mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py tests/test_gpu_cli.py tests/test_gpu_superkmers.py -x -q 2>&1 | tail -4
B="python bench.py --no-cpu-baseline --no-env"
$B --steps 3 --warmup 2 > gpurun_out/r3j_bench.log 2> gpurun_out/r3j_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3j_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], r['frac'], 'e2e', d['e2e'])"
