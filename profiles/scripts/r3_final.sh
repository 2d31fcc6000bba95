# round 2, final capture of session 3: full GPU suite, smoke, default bench + reference arm, launch list of one step (per-kernel time and DRAM bytes),
# full ncu captures of the hottest kernels at a reduced size
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3z_pytest.log; tail -3 gpurun_out/r3z_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
python bench.py > gpurun_out/r3z_bench.log 2> gpurun_out/r3z_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r3z_bench.log').read().strip().splitlines()[-1]); r=d['roofline']; print('cfg2', d['value'], d['ms_per_step'], r['frac'], r['stage_ms_per_step'], 'allocs', d['driver_allocs_in_timed_region'], 'e2e', d['e2e']['value'], 'env', d['env']['env_ms'], d['env']['reads_to_env_s'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('env_s'), d['clocks'])"
tail -2 gpurun_out/r3z_bench.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r3z_bench_ref.log 2>&1; cut -c1-300 gpurun_out/r3z_bench_ref.log
B="python bench.py --no-e2e --no-cpu-baseline --no-env"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r3z_launches.csv $B --steps 1 --warmup 0 > gpurun_out/r3z_launches.log 2>&1
python profiles/scripts/launch_summary.py gpurun_out/r3z_launches.csv gpurun_out/r3z_kernels.json | head -22
# full captures at 8 M reads (a process that holds > 100 GB makes every replayed pass save and restore device memory)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bin_agg' -s 1 -c 1 -o gpurun_out/r3z_agg $B --reads 8000000 --steps 1 --warmup 0 > gpurun_out/r3z_ncu1.log 2>&1; tail -1 gpurun_out/r3z_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bin_scatter2' -s 2 -c 1 -o gpurun_out/r3z_scatter2 $B --reads 8000000 --steps 1 --warmup 0 > gpurun_out/r3z_ncu2.log 2>&1; tail -1 gpurun_out/r3z_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_extract' -s 0 -c 1 -o gpurun_out/r3z_extract $B --reads 8000000 --steps 1 --warmup 0 > gpurun_out/r3z_ncu3.log 2>&1; tail -1 gpurun_out/r3z_ncu3.log
ls -la gpurun_out/r3z*.ncu-rep
