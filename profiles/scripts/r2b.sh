# round 2, capture "b": parity + A/B of the bulk-copy scatter + launch list of one step + environment-stage trace
mkdir -p gpurun_out
python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py tests/test_gpu_env.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2b_pytest.log; tail -3 gpurun_out/r2b_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline"
$B --no-env > gpurun_out/r2b_bench_tma1.log 2> gpurun_out/r2b_bench_tma1.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2b_bench_tma1.log').read().strip().splitlines()[-1]); print('tma1', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B --no-env --tune scatter_tma=0 > gpurun_out/r2b_bench_tma0.log 2> gpurun_out/r2b_bench_tma0.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2b_bench_tma0.log').read().strip().splitlines()[-1]); print('tma0', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
MC_TRACE=1 $B --steps 1 --warmup 1 > gpurun_out/r2b_bench_env.log 2> gpurun_out/r2b_env_trace.err; grep -v "agg group\|agg sample\|sub-bin" gpurun_out/r2b_env_trace.err | tail -60
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2b_launches.csv $B --no-env --steps 1 --warmup 0 > gpurun_out/r2b_launches.log 2>&1
python profiles/scripts/launch_summary.py gpurun_out/r2b_launches.csv gpurun_out/r2b_kernels.json | head -40
