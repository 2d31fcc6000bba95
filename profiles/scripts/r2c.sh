# round 2, capture "c": parser with 32 KiB tiles + per-terminator marks, warp-queued rehash; bench + launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2c_pytest.log; tail -3 gpurun_out/r2c_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
$B > gpurun_out/r2c_bench.log 2> gpurun_out/r2c_bench.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2c_bench.log').read().strip().splitlines()[-1]); print('r2c', d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])"
tail -3 gpurun_out/r2c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c_launches.csv $B --steps 1 --warmup 0 > gpurun_out/r2c_launches.log 2>&1
python profiles/scripts/launch_summary.py gpurun_out/r2c_launches.csv gpurun_out/r2c_kernels.json | head -24
