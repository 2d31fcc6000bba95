# round 2, capture "e": full GPU suite, default bench (cfg2), then cfg5 and cfg4 at N = 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2e_pytest.log; tail -3 gpurun_out/r2e_pytest.log
MC_TRACE=1 python bench.py > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2e_bench.log').read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'], d['env'], d['cpu_baseline'], d['e2e'])"
grep -v "agg group\|agg sample\|sub-bin" gpurun_out/r2e_bench.err | tail -16
python bench.py --workload cfg5 --steps 1 --warmup 1 --no-env > gpurun_out/r2e_bench_cfg5.log 2> gpurun_out/r2e_bench_cfg5.err; tail -c 3000 gpurun_out/r2e_bench_cfg5.log; tail -3 gpurun_out/r2e_bench_cfg5.err
python bench.py --workload cfg4 --steps 1 --warmup 1 > gpurun_out/r2e_bench_cfg4.log 2> gpurun_out/r2e_bench_cfg4.err; tail -c 3500 gpurun_out/r2e_bench_cfg4.log; tail -3 gpurun_out/r2e_bench_cfg4.err
