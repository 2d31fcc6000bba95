# round 2, capture "f": overlapped flush (scatter2 of group g+1 beside agg of group g) A/B + launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_count.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
for ov in 1 0; do
$B --tune flush_overlap=$ov > gpurun_out/r2f_bench_ov$ov.log 2> gpurun_out/r2f_bench_ov$ov.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2f_bench_ov$ov.log').read().strip().splitlines()[-1]); print('overlap $ov', d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])"
tail -2 gpurun_out/r2f_bench_ov$ov.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv $B --steps 1 --warmup 0 > gpurun_out/r2f_launches.log 2>&1
python profiles/scripts/launch_summary.py gpurun_out/r2f_launches.csv gpurun_out/r2f_kernels.json | head -24
