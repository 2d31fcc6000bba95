# round 2, capture "l": 16-byte pad granule with bulk copies + lazy CR-LF flags; parity + bench + launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2l_pytest.log; tail -3 gpurun_out/r2l_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-env --steps 8 --warmup 3"
$B > gpurun_out/r2l_bench.log 2> gpurun_out/r2l_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2l_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('r2l', d['ms_per_step'], r['frac'], 'stage sum', r['stage_sum_ms_per_step'], r['stage_ms_per_step'])"
tail -2 gpurun_out/r2l_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2l_launches.csv $B --steps 1 --warmup 0 > gpurun_out/r2l_launches.log 2>&1
python profiles/scripts/launch_summary.py gpurun_out/r2l_launches.csv gpurun_out/r2l_kernels.json | head -24
